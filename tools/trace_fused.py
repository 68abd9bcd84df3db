"""Timeline of block_fused_kernel's roles (CTA 0): builds a -DFNO_FZ_TRACE variant of the library into build/trace/,
runs one launch at B=256 and prints per-tile intervals.  Usage (GPU box): python tools/trace_fused.py [--build-only]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cfdbench_b200 import build as B
out_dir = os.path.join(ROOT, "cfdbench_b200", "build", "trace")
so = os.path.join(out_dir, "libtrace.so")
if not os.path.exists(so) or "--build-only" in sys.argv:
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for s in B.SOURCES:
        o = os.path.join(out_dir, s[:-3] + ".o")
        objs.append(o)
        subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, "-DFNO_FZ_TRACE", "-c", os.path.join(B.CSRC, s), "-o", o])
    subprocess.check_call([B._nvcc(), "-shared", "-o", so, *objs, "-lcudart"])
    print("built", so)
    if "--build-only" in sys.argv:
        sys.exit(0)
import numpy as np, torch
from cfdbench_b200 import _lib
_lib.LIB_PATH = so
lib = _lib.load()
lib.fno_debug_fused_trace.argtypes = [C.c_void_p]
batch = 256
trace = torch.zeros(4 * 256 * 8 + 148 * 4, dtype=torch.int64, device="cuda")
img = torch.randn(batch * 147456 // 4, device="cuda").view(torch.uint8)
x = torch.randn(batch, 32, 64, 64, device="cuda").bfloat16()
w0t = torch.randn(32, 32, device="cuda") / 6
bias = torch.randn(32, device="cuda")
out = torch.empty_like(x)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    if it == 2:
        assert lib.fno_debug_fused_trace(trace.data_ptr()) == 0
    _lib.check(lib.fno_block_fused(img.data_ptr(), x.data_ptr(), w0t.data_ptr(), bias.data_ptr(), out.data_ptr(), batch, st), "fused")
    torch.cuda.synchronize()
assert lib.fno_debug_fused_trace(None) == 0
def timed(label):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, z in ev:
        a.record()
        _lib.check(lib.fno_block_fused(img.data_ptr(), x.data_ptr(), w0t.data_ptr(), bias.data_ptr(), out.data_ptr(), batch, st), "fused")
        z.record()
    torch.cuda.synchronize()
    print(label, "median us", np.median([a.elapsed_time(z) * 1e3 for a, z in ev]).round(1))
lib.fno_debug_fused_knock.argtypes = [C.c_int]
for bits, label in ((0, "full kernel"), (1, "no conv MMAs"), (2, "no E MMAs"), (4, "no GEMM1 MMAs"), (8, "no converter stores"), (16, "no GELU"), (32, "no output stores"),
                    (3, "no tile MMAs"), (7, "no MMAs"), (48, "no GELU, no stores"), (15, "no MMAs, no conv stores"), (55, "no MMAs, GELU, stores"), (63, "skeleton only")):
    assert lib.fno_debug_fused_knock(bits) == 0
    timed("knock-out %2d %-26s" % (bits, label))
assert lib.fno_debug_fused_knock(0) == 0
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for a, z in ev:
    a.record()
    _lib.check(lib.fno_block_fused(img.data_ptr(), x.data_ptr(), w0t.data_ptr(), bias.data_ptr(), out.data_ptr(), batch, st), "fused")
    z.record()
torch.cuda.synchronize()
print("standalone kernel, trace off: us per launch", sorted(round(a.elapsed_time(z) * 1e3, 1) for a, z in ev))
cta = trace.cpu().numpy()[4 * 256 * 8:].reshape(148, 4)
pro = cta[:, 1] - cta[:, 0]
tot = cta[:, 2] - cta[:, 0]
print("per-CTA prologue cycles: min %d median %d max %d;  total cycles: min %d median %d max %d" % (pro.min(), np.median(pro), pro.max(), tot.min(), np.median(tot), tot.max()))
print("CTA totals by unit count: 4 units (CTA 0..67) median %d, 3 units (68..147) median %d" % (np.median(tot[:68]), np.median(tot[68:])))
print("kernel span (first entry to last end): %d cycles" % (cta[:, 2].max() - cta[:, 0].min()))
t = trace.cpu().numpy()[:4 * 256 * 8].reshape(4, 256, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
np.save(os.path.join(ROOT, "gpurun_out", "trace_fused.npy"), rel)
print("MMA thread per tile: start, x_ok, d2_ok, conv issued, bt_ok, all issued   (cycles since first event)")
for T in range(40):
    r = rel[2, T]
    if r[0] < 0: break
    print(f"T{T:3d} " + " ".join(f"{v:7d}" for v in r[:6]) + f"   | waits x {r[1]-r[0]:5d} d2 {r[2]-r[1]:5d} conv {r[3]-r[2]:5d} bt {r[4]-r[3]:5d} E {r[5]-r[4]:5d} | tile {r[5]-r[0]:6d}")
print("GEMM1 stages: wait start, y_ok, issued")
for i in range(64):
    r = rel[3, i]
    if r[0] >= 0: print(f"k{i//8} st{i%8}: " + " ".join(f"{v:7d}" for v in r[:3]), f"  wait {r[1]-r[0]} issue {r[2]-r[1]}")
print("converter (warp 0): T: wait_slot start, got slot, filled")
for T in range(40):
    r = rel[0, T]
    if r[3] < 0: break
    print(f"T{T:3d} d1wait {r[0]:7d} d1ok {r[1]:7d} ld {r[2]:7d} | slotwait {r[3]:7d} got {r[4]:7d} filled {r[5]:7d}  (wait {r[4]-r[3]}, fill {r[5]-r[4]})")
print("epilogue (warp 12): wait start, d2 full, tmem read, stored")
for T in range(40):
    r = rel[1, T]
    if r[0] < 0: break
    print(f"T{T:3d} " + " ".join(f"{v:7d}" for v in r[:4]) + f"  wait {r[1]-r[0]} ld {r[2]-r[1]} math+store {r[3]-r[2]}")
