"""EMD microbench: ours vs the reference's own kernels compiled for sm_100a (oracle/_ref/libemd_ref.so), same GPU.
Prints one JSON line.  python scripts/bench_emd.py [--b 64] [--n 1024]"""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from puzzlenet_b200 import emd_cuda


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--n", type=int, default=1024)
    a = ap.parse_args()
    b, n, m = a.b, a.n, a.n
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    x1 = (torch.randn(b, n, 3, generator=g) * 0.5).to(dev)
    x2 = (torch.randn(b, m, 3, generator=g) * 0.5).to(dev)
    gc = torch.ones(b, device=dev)
    match = emd_cuda.approxmatch_forward(x1, x2)
    res = {"b": b, "n": n, "m": m,
           "ours_ms": {"approxmatch": timed(lambda: emd_cuda.approxmatch_forward(x1, x2)),
                       "matchcost": timed(lambda: emd_cuda.matchcost_forward(x1, x2, match)),
                       "matchcost_backward": timed(lambda: emd_cuda.matchcost_backward(gc, x1, x2, match))}}
    so = os.path.join(ROOT, "oracle", "_ref", "libemd_ref.so")
    if os.path.isfile(so):
        ref = ctypes.CDLL(so)
        vp = ctypes.c_void_p
        rmatch = torch.empty(b, m, n, device=dev)
        temp = torch.empty(32 * (n + m) * 2, device=dev)
        cost = torch.empty(b, device=dev)
        g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, m, 3, device=dev)
        st = vp(torch.cuda.current_stream().cuda_stream)
        res["reference_kernels_ms"] = {
            "approxmatch": timed(lambda: ref.emd_ref_approxmatch(b, n, m, vp(x1.data_ptr()), vp(x2.data_ptr()), vp(rmatch.data_ptr()), vp(temp.data_ptr()), st)),
            "matchcost": timed(lambda: ref.emd_ref_matchcost(b, n, m, vp(x1.data_ptr()), vp(x2.data_ptr()), vp(rmatch.data_ptr()), vp(cost.data_ptr()), st)),
            "matchcost_backward": timed(lambda: ref.emd_ref_matchcost_grad(b, n, m, vp(gc.data_ptr()), vp(x1.data_ptr()), vp(x2.data_ptr()), vp(rmatch.data_ptr()), vp(g1.data_ptr()), vp(g2.data_ptr()), st)),
        }
        res["speedup"] = {k: round(res["reference_kernels_ms"][k] / res["ours_ms"][k], 2) for k in res["ours_ms"]}
        res["max_abs_match_diff"] = (match - rmatch).abs().max().item()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
