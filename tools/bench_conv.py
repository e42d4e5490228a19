"""Per-layer micro-benchmark of the convolution / BatchNorm kernels against the library (cuDNN / ATen) on the ResNet-18/CIFAR
shapes at B=128.  Every op is captured TEN times back to back in a CUDA graph and the graph replay is timed with CUDA events
(device time per launch; no Python / ctypes / tensor-map-encode time in the number, L2-warm as in the real step where the
activation was just produced by the previous layer).  Writes gpurun_out/conv_bench.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200.ops import conv as C  # noqa: E402
from draco_b200.ops import norm as NM  # noqa: E402

torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
dev = torch.device("cuda", 0)
REP = 10


def timeit(fn, replays=8):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm up on a side stream (autograd + capture dislike the legacy stream)
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(replays):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return round(s.elapsed_time(e) / (replays * REP) * 1e3, 2)          # us per launch


def cl(t):
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def wt(k, c, ks):
    return (torch.randn(k, c, ks, ks, device=dev) * 0.05).to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


rows = []
N = int(os.environ.get("BENCH_BATCH", "128"))
shapes = [(64, 64, 32, 3, 1), (128, 128, 16, 3, 1), (256, 256, 8, 3, 1), (512, 512, 4, 3, 1),
          (64, 128, 32, 3, 2), (128, 256, 16, 3, 2), (256, 512, 8, 3, 2),
          (64, 128, 32, 1, 2), (128, 256, 16, 1, 2), (256, 512, 8, 1, 2)]
for (c, k, hw, ks, st) in shapes:
    pad = ks // 2
    x = cl(torch.randn(N, c, hw, hw, device=dev))
    w = wt(k, c, ks)
    dy = cl(torch.randn(N, k, hw // st, hw // st, device=dev))
    args = (None, [st, st], [pad, pad], [1, 1], False, [0, 0], 1)
    halo = ks == 3 and st == 1 and C.halo_supported(hw, hw, c, k)
    req = C.BnStatRequest(1e-5, 0.1)
    if halo:
        f = lambda: C.conv3x3_halo(x, w)                                  # noqa: E731
        fs = lambda: C.conv3x3_halo(x, w, False, None, req)               # noqa: E731
        d = lambda: C.conv3x3_halo(dy, w, True)                           # noqa: E731
    else:
        f = lambda: C.convg_tcgen05(x, w, (hw, hw), st)                   # noqa: E731
        fs = lambda: C.convg_tcgen05(x, w, (hw, hw), st, False, None, req)  # noqa: E731
        d = lambda: C.convg_tcgen05(dy, w, (hw, hw), st, True)            # noqa: E731
    row = {"Cin": c, "Cout": k, "HW": hw, "ks": ks, "stride": st, "kernel": "halo" if halo else "tap",
           "fprop_us": timeit(f), "fprop_bnstats_us": timeit(fs),
           "cudnn_fprop_us": timeit(lambda: F.conv2d(x, w, stride=st, padding=pad)),
           "dgrad_us": timeit(d),
           "cudnn_dgrad_us": timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, *args, [True, False, False])),
           "wgrad_us": timeit(lambda: C.convg_wgrad_tcgen05(dy, x, ks, st)),
           "cudnn_wgrad_us": timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, *args, [False, True, False]))}
    if halo:
        row["tap_fprop_us"] = timeit(lambda: C.convg_tcgen05(x, w, (hw, hw), st))
        row["tap_wgrad_us"] = row["wgrad_us"]
        row["wgrad_us"] = timeit(lambda: C.conv3x3_halo_wgrad(dy, x))
    fl = 2.0 * N * (hw // st) ** 2 * k * c * ks * ks
    row["fprop_tflops"] = round(fl / row["fprop_us"] / 1e6, 1)
    rows.append(row)
    print(row, flush=True)

# stem
x = cl(torch.randn(N, 3, 32, 32, device=dev))
w = wt(64, 3, 3)
dy = cl(torch.randn(N, 64, 32, 32, device=dev))
args = (None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1)
req = C.BnStatRequest(1e-5, 0.1)
row = {"layer": "stem 3->64 32x32", "fprop_us": timeit(lambda: C.conv_stem_fprop(x, w)),
       "fprop_bnstats_us": timeit(lambda: C.conv_stem_fprop(x, w, None, req)),
       "cudnn_fprop_us": timeit(lambda: F.conv2d(x, w, padding=1)),
       "wgrad_us": timeit(lambda: C.conv_stem_wgrad(dy, x)),
       "cudnn_wgrad_us": timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, *args, [False, True, False]))}
rows.append(row)
print(row, flush=True)

# BatchNorm (+ReLU): forward with / without given statistics, backward; vs ATen batch_norm + relu
for (c, hw) in [(64, 32), (128, 16), (256, 8), (512, 4)]:
    x = cl(torch.randn(N, c, hw, hw, device=dev))
    gy = cl(torch.randn(N, c, hw, hw, device=dev))
    bn = NM.FusedBatchNorm2d(c).to(dev)
    mean, invstd = x.float().mean((0, 2, 3)), (x.float().var((0, 2, 3), unbiased=False) + 1e-5).rsqrt()
    xg = x.clone().requires_grad_(True)
    y = bn(xg, relu=True)

    def fwd_given():
        return NM._BnActFn.apply(x, None, bn.weight, bn.bias, None, None, 1e-5, 0.1, True, mean, invstd)

    def fwd_full():
        return NM._BnActFn.apply(x, None, bn.weight, bn.bias, None, None, 1e-5, 0.1, True, None, None)

    sy = y.detach()

    def bwd():                                           # the backward kernels directly (autograd's engine thread cannot be captured here)
        return NM._BnActFn.backward(ctx_stub, gy)

    class _Ctx:
        saved_tensors = (x, None, bn.weight, bn.bias, mean, invstd)
        relu, has_res = True, False
        link = None
    ctx_stub = _Ctx()

    row = {"layer": f"bn+relu C={c} {hw}x{hw}", "MB": round(x.numel() * 2 / 1e6, 1), "fwd_apply_only_us": timeit(fwd_given),
           "fwd_stats_apply_us": timeit(fwd_full), "bwd_us": timeit(bwd),
           "aten_fwd_us": timeit(lambda: F.relu(F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.1, 1e-5))),
           }
    rows.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/conv_bench.json", "w"), indent=1)
