"""CPU: the `attention_cutlass` extension keeps the reference's Python surface
(/root/reference/flash_attention_cutlass/csrc/attention_api.cpp:6-10, include/attention_api.h:10-11)."""
import pytest
import torch


@pytest.fixture(scope="module")
def mod(built):
    import attention_cutlass
    return attention_cutlass


def test_names_and_alias(mod):
    assert hasattr(mod, "flash_attention_v2_cutlass")
    assert hasattr(mod, "flash_attn_fwd")          # the name BASELINE.json uses
    assert mod.abi_version() == 2


def test_signature_is_positional_five_args(mod):
    doc = mod.flash_attention_v2_cutlass.__doc__
    # same pybind signature as the rebuilt reference module
    assert "arg0: torch.Tensor, arg1: torch.Tensor, arg2: torch.Tensor, arg3: bool, arg4:" in doc
    assert "-> list[torch.Tensor]" in doc
    q = torch.zeros(1, 1, 128, 64, dtype=torch.float16)
    with pytest.raises(TypeError):
        mod.flash_attention_v2_cutlass(q, q, q)               # 3-arg call is a TypeError, as in the reference


def test_cpu_tensor_rejected_like_reference(mod):
    q = torch.zeros(1, 1, 128, 64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="q must be a CUDA tensor"):
        mod.flash_attention_v2_cutlass(q, q, q, False, 1.0)


def test_import_does_not_need_gpu(mod):
    assert mod.launch_count() >= 0
