"""Per-worker compute microbenchmark: ResNet-18 fwd+bwd (B=128, bf16, channels-last) as a CUDA graph, fused-BN vs ATen-BN,
plus a torch.profiler kernel table of one eager iteration (real execution, warm L2 -- unlike ncu's cold-cache replays)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200 import JobConfig  # noqa: E402
from draco_b200.data import synthetic_dataset  # noqa: E402
from draco_b200.parallel.fused_engine import make_plan  # noqa: E402
from draco_b200.parallel.worker import WorkerCompute  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
net = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
out_dir = "gpurun_out"
os.makedirs(out_dir, exist_ok=True)
from draco_b200.ops import conv as _C  # noqa: E402
MODES = os.environ.get("WORKER_MODES", "fused,fused_wgrad_stream,aten").split(",")
for mode in MODES:
    # fused[_wgrad_stream][_nofork][_maskx][_bnfuse][_hiprio]: A/B switches of single optimisations, same process, same box
    os.environ["DRACO_BN"] = "aten" if mode == "aten" else "fused"
    os.environ["DRACO_CONV_FORK"] = "0" if "nofork" in mode else "1"
    os.environ["DRACO_BN_MASK"] = "x" if "maskx" in mode else "y"
    os.environ["DRACO_BN_BWD_FUSE"] = "1" if "bnfuse" in mode else "0"
    _C.WGRAD_SIDE_STREAM = "wgrad_stream" in mode
    cfg = JobConfig(network=net, dataset="Cifar10", approach="baseline", mode="normal", batch_size=128, num_workers=1,
                    dtype="bf16", synthetic_size=1024, transport="nvl").resolve(1)
    ds = synthetic_dataset("Cifar10", 1024)
    wc = WorkerCompute(cfg, dev, [1], make_plan(cfg, ds, None), ds)
    wc.stage_batches(1)
    for _ in range(3):
        wc.forward_backward(1, None)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    # "_hiprio": capture on a high-priority stream, so that the backward chain's kernels are dispatched ahead of the (default
    # priority) weight-gradient side stream whenever both have CTAs pending
    cap = torch.cuda.Stream(device=dev, priority=-1) if "hiprio" in mode else None
    with torch.cuda.graph(g, stream=cap):
        wc.forward_backward(1, None)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(30):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 30
    print(f"{net} {mode}: {ms:.3f} ms per fwd+bwd (graph replay), loss {wc.metrics[1][0].item():.4f}", flush=True)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        wc.forward_backward(1, None)
        torch.cuda.synchronize()
    tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90)
    with open(os.path.join(out_dir, f"worker_profile_{net}_{mode}{os.environ.get('PROFILE_TAG', '')}.txt"), "w") as fh:
        fh.write(f"{net} {mode}: {ms:.3f} ms per fwd+bwd (graph replay)\n\n" + tab)
    del wc, g
