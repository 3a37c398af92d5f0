"""Kernel times (library events) of the fused-prologue entry points next to the plain ones at BASELINE config 2,
direct C-ABI calls:  python tools/dev_fused.py [dist]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from weed_instance_segmentation_b200 import _cabi, synth  # noqa: E402

H, D, L, P = bench.H, bench.D, bench.L, bench.P


def main():
    dist = sys.argv[1] if len(sys.argv) > 1 else "init"
    dev = torch.device("cuda", 0)
    pr = bench.Problem(dist, "bf16", dev, seed=0)
    lib, c = pr.lib, _cabi
    B, S = pr.batch, pr.S
    g = torch.Generator(device="cuda").manual_seed(1)
    off = (synth.init_offsets(H, L, P).to(dev)[None, None]
           + 0.5 * torch.randn(B, S, H, L, P, 2, device=dev, generator=g)).bfloat16().contiguous()
    logits = torch.randn(B, S, H, L * P, device=dev, generator=g).bfloat16()
    goff, glog = torch.empty_like(off), torch.empty_like(logits)
    p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    rows = {"plain fwd": [], "plain bwd": [], "fused fwd": [], "fused bwd": []}
    for _ in range(9):
        pr.fwd(pr.pdesc)
        torch.cuda.synchronize()
        rows["plain fwd"].append(c.profile_ms(c.PROF_FWD))
        pr.bwd(pr.pdesc)
        torch.cuda.synchronize()
        rows["plain bwd"].append(c.profile_ms(c.PROF_BWD_MAIN))
        c.check(lib.msda_b200_forward_fused(pr.pdesc, p(pr.value), p(off), p(logits), None, p(pr.out), None, p(pr.order),
                                            pr.stream))
        torch.cuda.synchronize()
        rows["fused fwd"].append(c.profile_ms(c.PROF_FWD))
        c.check(lib.msda_b200_backward_fused(pr.pdesc, p(pr.value), p(off), p(logits), None, p(pr.go), p(pr.gv), p(goff),
                                             p(glog), p(pr.ws), pr.nws, p(pr.order), pr.stream))
        torch.cuda.synchronize()
        rows["fused bwd"].append(c.profile_ms(c.PROF_BWD_MAIN))
    print(f"config 2 / {dist} / bf16, kernel ms (median of 9): " +
          " | ".join(f"{k} {sorted(v)[4]:.3f}" for k, v in rows.items()), flush=True)


if __name__ == "__main__":
    main()
