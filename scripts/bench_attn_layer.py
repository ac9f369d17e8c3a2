"""Fused offset-attention layer (attention_layer_tc.cu) alone: time per launch at C clouds and the kernel-internal
timeline of CTA 0 (pz_profile_attention_timeline).  Goes through pz_offset_attention(precision=bf16), whose
conversion launches are timed separately and subtracted.

    python scripts/bench_attn_layer.py [--clouds 128] [--iters 50]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlenet_b200 import _lib  # noqa: E402

EPI = ["start", "x back for r", "q|k acc ready", "q|k written", "v^T acc ready", "v^T written", "S ready", "P0 written",
       "P1 max done", "P v 0 done", "P1 written", "P v 1 done", "r written", "out acc ready", "TMA stores have read"]
MMA = {32: "mma start", 33: "phase 1 issued", 34: "v^T ch0 issued", 35: "q|k drained", 36: "v^T issued", 37: "v^T drained",
       38: "P0 ready", 39: "P1 ready", 40: "r ready", 41: "out issued"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=128)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--predict5", action="store_true", help="timeline of the 4th layer inside a B=64 predict5 forward")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    tl = torch.zeros(64 + 2 * 4096, device=dev, dtype=torch.int64)
    if a.predict5:
        import types
        from puzzlenet_b200.model5_b import TouchedRegraster
        from puzzlenet_b200.weights import make_batch, synthetic_pairs, synthetic_state_dict
        model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
        model.load_state_dict(synthetic_state_dict(0), strict=True)
        model.to(dev).eval()
        model.precision = "bf16"
        fpc, mrpc = synthetic_pairs(64, seed=64)
        batch = make_batch(fpc.to(dev), mrpc.to(dev))
        for _ in range(3):
            model.predict5(batch, 0)
        torch.cuda.synchronize()
        _lib.call("pz_profile_attention_timeline", tl.data_ptr(), tl.numel())
        model.predict5(batch, 0)
        torch.cuda.synchronize()
        _lib.call("pz_profile_attention_timeline", None, 0)
        t = tl.cpu().tolist()
        report(t, 128, None)
        report_gather(t)
        return
    B, L, C = a.clouds, 256, 256
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, L, C, generator=g) * 0.5).to(dev)
    W = [(torch.randn(n, C, generator=g) / 16).to(dev) for n in (64, 64, 256, 256)]
    bs = [torch.randn(n, generator=g).to(dev) * 0.1 for n in (64, 64, 256, 256)]
    out = torch.empty_like(x)
    ws_bytes = lib.pz_offset_attention_workspace_bytes(B, L, C)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    args = [x.data_ptr(), W[0].data_ptr(), bs[0].data_ptr(), W[1].data_ptr(), bs[1].data_ptr(), W[2].data_ptr(),
            bs[2].data_ptr(), W[3].data_ptr(), bs[3].data_ptr(), B, L, C, _lib.PZ_PREC_BF16, out.data_ptr(), None,
            ws.data_ptr(), ws_bytes, _lib.stream_ptr()]

    def run():
        _lib.call("pz_offset_attention", *args)

    for _ in range(5):
        run()
    torch.cuda.synchronize()
    # CUDA graph of `iters` calls, so launch gaps do not count; kernel time from a graph of the conversions alone is
    # not separable here, so report the whole call and the in-kernel clock span
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms_call = e0.elapsed_time(e1) / a.iters
    _lib.call("pz_profile_attention_timeline", tl.data_ptr(), tl.numel())
    run()
    torch.cuda.synchronize()
    _lib.call("pz_profile_attention_timeline", None, 0)
    report(tl.cpu().tolist(), B, ms_call)


def report_gather(t):
    """CTA 0 of the stage-1 gather GEMM (tc_gemm_kernel<256,1,true,true>): producer group 0 (slots 1024 + 4 n: job
    start, data landed, Q done, transformed + published), MMA issuer (1280 + job: operands ready), epilogue (1408 + 2 tile:
    accumulator ready, tile done)."""
    mhz = 1965.0
    t0 = t[1024]
    if t0 == 0:
        return
    print("gather GEMM (stage 1), CTA 0; us since producer group 0 started job 0")
    for n in range(10):
        a = [(t[1024 + 4 * n + k] - t0) / mhz for k in range(4)]
        print(f"  group-0 job {2 * n:2d}: start {a[0]:7.2f}  landed {a[1]:7.2f}  Q {a[2]:7.2f}  published {a[3]:7.2f}")
    print("  MMA: operands of job n ready at", [round((t[1280 + n] - t0) / mhz, 2) for n in range(20)])
    print("  epilogue (acc ready, done) per tile:", [(round((t[1408 + 2 * n] - t0) / mhz, 2), round((t[1409 + 2 * n] - t0) / mhz, 2))
                                                      for n in range(8)])


def report(t, B, ms_call):
    t0 = t[15]
    mhz = 1965.0
    rows = [(t[i] - t0, "epi", EPI[i]) for i in range(len(EPI))] + [(t[16] - t0, "epi", "x back for out"), (t[17] - t0, "epi", "out tiles formed")] + [(t[k] - t0, "mma", v) for k, v in MMA.items()]
    rows.sort()
    for cyc, who, what in rows:
        print(f"{cyc:8d} cyc {cyc / mhz:7.2f} us  {who}  {what}")
    flop = B * (2 * 256 * 256 * 384 + 2 * 256 * 256 * 64 + 2 * 256 * 256 * 256 + 2 * 256 * 256 * 256)
    span_us = (t[14] - t0) / mhz
    print(f"entry -> start {(t[0] - t0) / mhz:.2f} us")
    ent = [t[64 + 2 * c] for c in range(B)]
    ext = [t[65 + 2 * c] for c in range(B)]
    e0 = min(ent)
    spans = sorted(x - e for e, x in zip(ent, ext))
    print(f"all CTAs (globaltimer): first entry -> last entry {(max(ent) - e0) / 1e3:.2f} us, first entry -> last exit "
          f"{(max(ext) - e0) / 1e3:.2f} us, CTA span min/median/max {spans[0] / 1e3:.2f}/{spans[len(spans) // 2] / 1e3:.2f}/"
          f"{spans[-1] / 1e3:.2f} us")
    print(json.dumps({"clouds": B, "ms_per_call_incl_conversions": ms_call, "cta0_span_us": span_us,
                      "tflops_if_all_ctas_like_cta0": flop / B * min(B, 148) / (span_us * 1e-6) / 1e12}))


if __name__ == "__main__":
    main()
