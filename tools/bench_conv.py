"""3x3 convolution: tcgen05 TMA-patch implicit GEMM (ours) vs cuDNN, ResNet-18/CIFAR shapes at B=128 (CUDA events, warm L2 as
in the real step: the activation was just produced by the previous layer)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200.ops.conv import conv3x3_tcgen05, conv3x3_wgrad_tcgen05  # noqa: E402

torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
dev = torch.device("cuda", 0)


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3          # us


rows = []
for (n, c, k, hw) in [(128, 64, 64, 32), (128, 128, 128, 16), (128, 256, 256, 8), (128, 512, 512, 4), (128, 64, 128, 16)]:
    x = torch.randn(n, c, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(k, c, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(n, k, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    fl = 2.0 * n * hw * hw * k * c * 9
    t_f = timeit(lambda: conv3x3_tcgen05(x, w))
    t_fc = timeit(lambda: F.conv2d(x, w, padding=1))
    t_d = timeit(lambda: conv3x3_tcgen05(dy, w, True))
    t_dc = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]))
    t_w = timeit(lambda: conv3x3_wgrad_tcgen05(dy, x))
    t_wc = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]))
    rows.append({"wgrad_us": t_w, "cudnn_wgrad_us": t_wc, "N": n, "Cin": c, "Cout": k, "HW": hw, "fprop_us": t_f, "cudnn_fprop_us": t_fc, "dgrad_us": t_d, "cudnn_dgrad_us": t_dc,
                 "fprop_tflops": fl / t_f / 1e6, "cudnn_fprop_tflops": fl / t_fc / 1e6})
    print(rows[-1], flush=True)
if os.environ.get("DRACO_EXPERIMENTAL", "0") == "1":
    # halo-reuse kernels on the 64 -> 64 layer1 shape vs the per-tap kernel and cuDNN
    from draco_b200.ops.conv import conv3x3_halo  # noqa: E402
    try:
        x = torch.randn(128, 64, 32, 32, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        row = {"shape": "128x64x32x32 -> 64", "halo_fprop_us": timeit(lambda: conv3x3_halo(x, w)),
               "halo_dgrad_us": timeit(lambda: conv3x3_halo(x, w, True)), "tap_fprop_us": timeit(lambda: conv3x3_tcgen05(x, w)),
               "cudnn_fprop_us": timeit(lambda: F.conv2d(x, w, padding=1))}
        rows.append(row)
        print(row, flush=True)
    except Exception as e:  # noqa: BLE001  (keep the rest of the bench alive)
        print("halo bench failed:", e, flush=True)
    # strided / 1x1 layers of ResNet-18 on the tap-table kernels vs cuDNN (the stride-2 dgrads are cuDNN's slowest kernels here)
    from draco_b200.ops.conv import convg_tcgen05, convg_wgrad_tcgen05  # noqa: E402
    for (n, c, k, hw, ks) in [(128, 64, 128, 32, 3), (128, 128, 256, 16, 3), (128, 256, 512, 8, 3), (128, 64, 128, 32, 1),
                              (128, 128, 256, 16, 1), (128, 256, 512, 8, 1)]:
        pad = ks // 2
        x = torch.randn(n, c, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(k, c, ks, ks, device=dev) * 0.05).to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        dy = torch.randn(n, k, hw // 2, hw // 2, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        args = (None, [2, 2], [pad, pad], [1, 1], False, [0, 0], 1)
        row = {"N": n, "Cin": c, "Cout": k, "HW": hw, "ks": ks, "stride": 2,
               "fprop_us": timeit(lambda: convg_tcgen05(x, w, (hw, hw), 2)),
               "cudnn_fprop_us": timeit(lambda: F.conv2d(x, w, stride=2, padding=pad)),
               "dgrad_us": timeit(lambda: convg_tcgen05(dy, w, (hw, hw), 2, True)),
               "cudnn_dgrad_us": timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, *args, [True, False, False])),
               "wgrad_us": timeit(lambda: convg_wgrad_tcgen05(dy, x, ks, 2)),
               "cudnn_wgrad_us": timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, *args, [False, True, False]))}
        rows.append(row)
        print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/conv_bench.json", "w"), indent=1)
