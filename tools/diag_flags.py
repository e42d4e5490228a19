"""Diagnostic: same-GPU cross-stream flag handshake (spin kernel on one stream, producer on another)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200.ops import kernels as K
from draco_b200.models import build_model
from draco_b200.parallel.arena import ArenaLayout

dev = torch.device("cuda", 0)
L = ArenaLayout.from_model(build_model("LeNet"), bf16=False, channels_last=True)
L.tile_view(dev)                                   # upload tables before anything spins
g32 = torch.randn(L.total, device=dev)
dst = torch.zeros(L.total, device=dev)
step = torch.full((1,), 9, dtype=torch.int64, device=dev)
cnt = torch.zeros(4, dtype=torch.int32, device=dev)
flags = torch.zeros(8, dtype=torch.int64, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
for variant in ("preuploaded", "set_flags_only"):
    flags.zero_(); err.zero_(); torch.cuda.synchronize()
    t0 = time.time()
    with torch.cuda.stream(s2):
        K.wait_flags([flags[0:1]], step, 0, err, timeout_s=5.0)
    t1 = time.time()
    with torch.cuda.stream(s1):
        torch.cuda._sleep(2_000_000)
        if variant == "preuploaded":
            K.push_encode(L, [g32], [None], dst, step_ptr=step, worker=0, done_counter=cnt[0:1], flag=flags[0:1])
        else:
            K.set_flags([flags[0:1]], step, 0)
    t2 = time.time()
    torch.cuda.synchronize()
    t3 = time.time()
    print(variant, "launch_wait %.4f launch_prod %.4f sync %.4f" % (t1 - t0, t2 - t1, t3 - t2), "err", err.item(), "flag", flags[0].item(), flush=True)
