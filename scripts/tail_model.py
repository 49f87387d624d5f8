"""List-scheduling model of the grid: how much time does the launch order of work items cost on causal problems?
CTAs are handed to SMs in blockIdx order as SMs free up (one CTA per SM).  Item cost = KV tile iterations x PERIOD +
OVERHEAD cycles (3250 and 6100).  Compares the shipped order ((b,h)-major, heaviest pair first
inside a head) with alternatives.  CPU only.   python scripts/tail_model.py
"""
import heapq
import sys

PERIOD, OVERHEAD, SMS = 3250, 6100, 148


def makespan(costs):
    sms = [0] * SMS
    heapq.heapify(sms)
    end = 0
    for c in costs:
        t = heapq.heappop(sms) + c
        end = max(end, t)
        heapq.heappush(sms, t)
    return end


def items(BH, S, causal):
    npairs = (S + 255) // 256
    nkv = (S + 127) // 128
    per_head = []
    for pr in range(npairs - 1, -1, -1):                  # heaviest first inside a head
        n = min(nkv, (pr * 256 + 128) // 128 + 1) if causal else nkv
        per_head.append(n * PERIOD + OVERHEAD)
    return per_head


def report(name, BH, S, causal=True):
    ph = items(BH, S, causal)
    total = sum(ph) * BH
    ideal = total / SMS
    shipped = makespan([c for _ in range(BH) for c in ph])
    out = [f"{name}: ideal {ideal / 1e3:7.0f}k cycles; shipped order {shipped / 1e3:7.0f}k (+{100 * (shipped / ideal - 1):4.1f} %)"]
    for G in (8, 16, 32, 64, BH):
        if G > BH:
            continue
        order = []
        for c0 in range(0, BH, G):                        # chunks of G heads (K/V working set G x 2 S D 2 bytes)
            g = min(G, BH - c0)
            for c in ph:                                  # pair-major inside the chunk: all heavy items first
                order += [c] * g
        m = makespan(order)
        out.append(f"    chunks of {G:3d} heads, pair-major inside: {m / 1e3:7.0f}k (+{100 * (m / ideal - 1):4.1f} %)")
    print("\n".join(out))


if __name__ == "__main__":
    report("cfg3  B4 H32 S4096  causal", 128, 4096)
    report("cfg5 shard B8 H32 S4096 causal", 256, 4096)
    report("cfg5  B64 H32 S4096 causal", 2048, 4096)
    report("cfg4  B1 H32 S16384 causal", 32, 16384)
    report("B4 H32 S4096 non-causal", 128, 4096, causal=False)
