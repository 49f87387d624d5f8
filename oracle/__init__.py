"""CPU oracle for the attention-forward hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product path (tiny-flash-attention_b200/, the C ABI, `attention_cutlass`) never does.
"""
from .oracle import (  # noqa: F401
    attn_exact,
    build,
    flash_v2_blocks,
    load_ref_kernels,
    numpy_attention,
    num_threads,
    rowwise_online,
)
