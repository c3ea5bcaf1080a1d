"""Time the tcgen05 GEMM on the encoder's shapes through the C ABI (suta_op_gemm), next to cuBLAS (torch.matmul) on the
same operands.  CUDA events on the launching stream, 3 warm-ups, rotating operand sets larger than L2.
Usage: python tools/gemm_bench.py [M]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]
import torch  # noqa: E402

from suta_b200 import _lib  # noqa: E402
from suta_b200._lib import check  # noqa: E402

lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20096
H, I = 768, 3072


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


FILTER = os.environ.get("GEMM_FILTER")


def run(name, N, K, bias=False, res=False, act=0, f32=False, bf16=True, acc=False, reps=20, nset=4, noaux=False):
    if FILTER and FILTER not in name:
        return
    sets = []
    for _ in range(nset):
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        d = dict(a=a, b=b,
                 o32=torch.empty(M, N, device="cuda") if f32 else None,
                 o16=torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if bf16 else None,
                 bias=torch.randn(N, device="cuda") if bias else None,
                 res=torch.randn(M, N, device="cuda") if res else None,
                 aux=torch.randn(M, N, device="cuda").bfloat16() if act else None)
        sets.append(d)
    st = torch.cuda.current_stream().cuda_stream

    def ours(d):
        check(lib.suta_op_gemm(P(d["a"]), M, K, P(d["b"]), N, K, M, N, K, P(d["o32"]), P(d["o16"]), N, P(d["bias"]),
                               P(d["res"]), N, act | (4 if acc else 0), P(d["aux"]) if act == 2 else None, P(d["aux"]) if act == 1 and not noaux else None, N, st))

    def cublas(d):
        torch.matmul(d["a"], d["b"].t())

    out = {}
    for tag, fn in (("ours", ours), ("cublas", cublas)):
        for i in range(3):
            fn(sets[i % nset])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(sets[i % nset])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[tag + "_us"] = round(ms * 1e3, 1)
        out[tag + "_tflops"] = round(2.0 * M * N * K / (ms * 1e-3) / 1e12, 1)
    print(f"{name:28s} M={M} N={N} K={K} " + json.dumps(out), flush=True)
    if os.environ.get("SUTA_GEMM_TRACE"):
        cap = 16
        tr = torch.zeros(cap, 8, dtype=torch.int64, device="cuda")
        lib.suta_debug_set_gemm_trace(P(tr), cap)
        ours(sets[0])
        torch.cuda.synchronize()
        lib.suta_debug_set_gemm_trace(None, 0)
        t = tr.cpu().numpy()
        t0 = t[0, 7]
        print("   iter: prod_start | mma_wait mma_start mma_issued | epiA_start epiA_end | epiB_start epiB_end   (cycles since first producer event)")
        for i in range(cap):
            if t[i, 1] == 0:
                break
            r = [int(t[i, k] - t0) if t[i, k] else -1 for k in (7, 0, 1, 2, 3, 4, 5, 6)]
            print(f"   {i:3d}: {r[0]:8d} | {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} | {r[6]:8d} {r[7]:8d}")


print("device", torch.cuda.get_device_name(0))
run("qkv  (+bias, bf16)", 3 * H, H, bias=True)
run("oproj(+bias+res, f32)", H, H, bias=True, res=True, f32=True, bf16=False)
run("oproj(+bias, f32 +=)", H, H, bias=True, acc=True, f32=True, bf16=False)
run("ffn1 (+bias, gelu, aux)", I, H, bias=True, act=1)
run("ffn1 (+bias, gelu, no aux)", I, H, bias=True, act=1, noaux=True)
run("ffn1 (+bias only)", I, H, bias=True)
run("ffn1 (gelu, aux, no bias)", I, H, act=1)
run("ffn2 (+bias, f32 +=)", H, I, bias=True, acc=True, f32=True, bf16=False)
run("ffn2 dgrad (x aux)", I, H, act=2)
run("ffn1 dgrad (f32 +=)", H, I, acc=True, f32=True, bf16=False)
run("oproj dgrad (bf16)", H, H)
run("qkv dgrad (f32 +=)", H, 3 * H, acc=True, f32=True, bf16=False)
run("conv k3 (gelu, bf16)", 512, 1536, act=1)
run("plain 4096^2-ish", 4096, 4096)
