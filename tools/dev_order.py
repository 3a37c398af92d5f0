"""Tuning run on a GPU box: forward (and backward) kernel time at BASELINE config 2 / 3 against the ORDER in which
queries are handed to thread blocks: 8x16 patches walked row-major (functional.query_order_2d today) or in sub-patches
(a forward block takes 32 consecutive queries of the order: a 2x16 strip today, a 4x8 patch with sub = 4x8).

    python tools/dev_order.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def order_2d(shapes, tile, sub, device):
    th, tw = tile
    sh, sw = sub
    parts, start = [], 0
    for h, w in shapes:
        y = torch.arange(h).view(h, 1).expand(h, w)
        x = torch.arange(w).view(1, w).expand(h, w)
        tiles_x = (w + tw - 1) // tw
        yy, xx = y % th, x % tw
        subs_x = (tw + sw - 1) // sw
        rank = (((y // th) * tiles_x + (x // tw)) * (th * tw)
                + ((yy // sh) * subs_x + (xx // sw)) * (sh * sw) + (yy % sh) * sw + (xx % sw))
        parts.append(start + torch.argsort(rank.reshape(-1), stable=True))
        start += h * w
    return torch.cat(parts).to(torch.int32).to(device)


def main():
    dev = torch.device("cuda", 0)
    c3 = [(31, 41), (61, 81), (121, 162)]
    for tag, shapes, batch in (("config2", None, None), ("config3", c3, 16)):
        for dist in ("init", "trained"):
            pr = bench.Problem(dist, "bf16", dev, seed=0, shapes=shapes, batch=batch)
            base = pr.order
            line = f"{tag}/{dist:8s}"
            for name, tile, sub in (("8x16 row-major", (8, 16), (8, 16)), ("sub 4x8", (8, 16), (4, 8)), ("sub 2x16", (8, 16), (2, 16)),
                                    ("sub 8x4", (8, 16), (8, 4)), ("sub 4x4", (8, 16), (4, 4)), ("sub 2x8", (8, 16), (2, 8)),
                                    ("16x8 sub 4x8", (16, 8), (4, 8))):
                pr.order = order_2d(pr.shapes, tile, sub, dev)
                ts = sorted((bench.profile_kernels(pr, 1) for _ in range(7)), key=lambda k: k["fwd"])
                tb = sorted(t["bwd_main"] for t in ts)
                line += f" | {name}: fwd {ts[3]['fwd']:.3f} bwd {tb[3]:.3f}"
            pr.order = base
            print(line, flush=True)
            del pr
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
