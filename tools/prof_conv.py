"""Launch each tcgen05 convolution kernel (fprop / dgrad / wgrad) a few times on two ResNet-18 layer shapes -- the ncu target
(`tools/gpu_ci.sh ncu_conv`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200.ops.conv import conv3x3_tcgen05, conv3x3_wgrad_tcgen05  # noqa: E402

dev = torch.device("cuda", 0)
for (n, c, k, hw) in [(128, 64, 64, 32), (128, 256, 256, 8)]:
    x = torch.randn(n, c, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(k, c, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(n, k, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        conv3x3_tcgen05(x, w)
        conv3x3_tcgen05(dy, w, True)
        conv3x3_wgrad_tcgen05(dy, x)
    torch.cuda.synchronize()
print("done")
