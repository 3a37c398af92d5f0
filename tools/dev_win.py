"""Development check of the window-staged kernels (csrc/msda_win.cu) against the per-corner kernels of msda_b200.cu:
same inputs through both paths, difference and CUDA-event timing. Run on a B200:  python tools/dev_win.py [fwd|bwd]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from weed_instance_segmentation_b200 import _cabi, build, functional  # noqa: E402
from weed_instance_segmentation_b200.synth import msda_inputs  # noqa: E402

os.environ["MSDA_B200_WINDOW"] = "1"  # the window kernels are opt-in
build.build()
lib = _cabi.load()
H, D, P = 8, 32, 4
CASES = [
    ("c2/init", 8, [(32, 32), (64, 64), (128, 128)], "init"),
    ("c2/trained", 8, [(32, 32), (64, 64), (128, 128)], "trained"),
    ("c2/adversarial", 8, [(32, 32), (64, 64), (128, 128)], "adversarial"),
    ("c3/init", 4, [(31, 41), (61, 81), (121, 162)], "init"),
    ("c5/init", 2, [(64, 64), (128, 128), (256, 256)], "init"),
    ("odd/trained", 3, [(5, 7), (9, 4)], "trained"),
    ("tiny/adversarial", 2, [(1, 1), (1, 9), (7, 1), (4, 4)], "adversarial"),
]


def ptr(t):
    return t.data_ptr() if t is not None else None


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    for tag, B, shapes, dist in CASES:
        L = len(shapes)
        x = msda_inputs(B, shapes, num_heads=H, head_dim=D, num_points=P, dist=dist, seed=3, device="cuda",
                        value_dtype=torch.bfloat16)
        value, loc, attn, go = x["value"], x["sampling_locations"], x["attention_weights"], x["grad_out"]
        S = value.shape[1]
        lsi = x["level_start_index"].tolist()
        order = functional.query_order_2d(shapes, functional._TILE, "cuda")
        sched = functional.pyramid_schedule(shapes, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        d_old, k1 = _cabi.make_desc(B, S, S, H, D, L, P, _cabi.BF16, _cabi.BF16, shapes, lsi, _cabi.FLAG_NO_WINDOW)
        d_new, k2 = _cabi.make_desc(B, S, S, H, D, L, P, _cabi.BF16, _cabi.BF16, shapes, lsi, 0, sched)
        out_old, out_new = torch.empty_like(go), torch.full_like(go, float("nan"))

        def f_old():
            _cabi.check(lib.msda_b200_forward(d_old, ptr(value), ptr(loc), ptr(attn), ptr(out_old), ptr(order), stream))

        def f_new():
            _cabi.check(lib.msda_b200_forward(d_new, ptr(value), ptr(loc), ptr(attn), ptr(out_new), ptr(sched.order), stream))

        if what == "bwd":  # backward timing through the C ABI
            gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
            nws = int(lib.msda_b200_backward_workspace_bytes(d_old))
            ws = torch.empty(nws, dtype=torch.uint8, device="cuda")

            def f_bwd():
                _cabi.check(lib.msda_b200_backward(d_old, ptr(value), ptr(loc), ptr(attn), ptr(go), ptr(gv), ptr(gl), ptr(ga),
                                                   ptr(ws), nws, ptr(order), stream))
            print(f"{tag:18s} backward {timed(f_bwd):.3f} ms   forward {timed(f_old):.3f} ms", flush=True)
            continue
        if what == "onebwd":  # a few launches of the backward on config 2 / init, for an ncu capture
            gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
            nws = int(lib.msda_b200_backward_workspace_bytes(d_old))
            ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
            for _ in range(3):
                _cabi.check(lib.msda_b200_backward(d_old, ptr(value), ptr(loc), ptr(attn), ptr(go), ptr(gv), ptr(gl), ptr(ga),
                                                   ptr(ws), nws, ptr(order), stream))
            torch.cuda.synchronize()
            return
        if what == "oneold":  # a few launches of the per-corner forward on config 2 / init, for an ncu capture
            for _ in range(3):
                f_old()
            torch.cuda.synchronize()
            return
        if what == "one":  # a few launches of the window forward on config 2 / init, for an ncu capture
            for _ in range(3):
                f_new()
            torch.cuda.synchronize()
            return
        if what == "phases":
            # clock64 stamps of thread 0 in the first 512 blocks (csrc/msda_win.cu dbg_mark)
            import ctypes
            dbg = torch.zeros(512 * 12, dtype=torch.int64, device="cuda")
            lib.msda_b200_internal_win_debug.argtypes = [ctypes.c_void_p]
            for nt in ("256",):
                os.environ["MSDA_B200_WIN_NT"] = nt
                lib.msda_b200_internal_win_debug(dbg.data_ptr())
                for _ in range(3):
                    f_new()
                torch.cuda.synchronize()
                lib.msda_b200_internal_win_debug(None)
                t = dbg.view(512, 12).cpu().double()
                names = {1: "pass1", 2: "tma-issue+desc", 3: "tma-wait", 11: "gather+epilogue"}
                prev, parts = 0, []
                for slot in (1, 2, 3, 11):
                    parts.append(f"{names[slot]}={(t[:, slot] - t[:, prev]).mean().item():.0f}")
                    prev = slot
                print(f"{tag:18s} NT={nt} cycles/block total={(t[:, 11] - t[:, 0]).mean().item():.0f} :: " + " ".join(parts),
                      flush=True)
            os.environ.pop("MSDA_B200_WIN_NT")
            continue
        if what == "pair":  # msda_fwd_kernel (one head per block) vs msda_fwd_pair_kernel (an even and an odd head)
            os.environ["MSDA_B200_FWD_NO_PAIR"] = "1"
            f_old()
            torch.cuda.synchronize()
            ref = out_old.clone()
            line = f"{tag:18s} one-head {timed(f_old):.3f} ms"
            os.environ.pop("MSDA_B200_FWD_NO_PAIR")
            for nt, qpg in (("128", "2"), ("128", "1"), ("128", "4"), ("256", "1"), ("256", "2")):
                os.environ["MSDA_B200_FWD_NT"], os.environ["MSDA_B200_FWD_QPG"] = nt, qpg
                out_old.fill_(float("nan"))
                f_old()
                torch.cuda.synchronize()
                line += f" | pair nt{nt} qpg{qpg} {timed(f_old):.3f} ms equal={torch.equal(out_old, ref)}"
            os.environ.pop("MSDA_B200_FWD_NT"), os.environ.pop("MSDA_B200_FWD_QPG")
            print(line, flush=True)
            continue
        if what == "fwd":
            f_old()
            line = f"{tag:18s} tiles={sched.num_tiles} max_tile={sched.max_tile} old {timed(f_old):.3f} ms"
            for name, env in (("cuda-core", {"MSDA_B200_WIN_KERNEL": "0"}), ("mma", {}), ("mma/hpb1", {"MSDA_B200_WIN_HPB": "1"}),
                              ("mma/hpb2", {"MSDA_B200_WIN_HPB": "2"}), ("mma/hpb8", {"MSDA_B200_WIN_HPB": "8"})):
                os.environ.update(env)
                out_new.fill_(float("nan"))
                f_new()
                torch.cuda.synchronize()
                line += f" | {name} {timed(f_new):.3f} ms rel={rel(out_new, out_old):.2e}"
                for key in env:
                    os.environ.pop(key)
            print(line, flush=True)


if __name__ == "__main__":
    main()
