"""CPU tests: cluster tool, bucket planning, op fallbacks (Conv2d / FusedBatchNorm2d / Linear on CPU are plain ATen)."""
import pytest
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from draco_b200.cli import cluster
from draco_b200.models import build_model
from draco_b200.ops.conv import Conv2d
from draco_b200.ops.linear import Linear
from draco_b200.ops.norm import FusedBatchNorm2d
from draco_b200.parallel.arena import ArenaLayout
from draco_b200.parallel.worker import plan_buckets


def test_cfg_self_interpolation_and_hosts(tmp_path):
    cfg = cluster.Cfg({"name": "job", "ssh_user": "me", "remote_dir": "/home/%(ssh_user)s/%(name)s",
                       "train_dir": "%(remote_dir)s/models/", "nodes": ["10.0.0.1", "10.0.0.2"]})
    assert cfg["remote_dir"] == "/home/me/job" and cfg["train_dir"] == "/home/me/job/models/"
    files = cluster.get_hosts(cfg, str(tmp_path))
    assert files["hosts_address"] == "10.0.0.1\n10.0.0.2\n" and "node1" in files["hosts"]
    assert (tmp_path / "hosts_alias").read_text().split() == ["node0", "node1"]
    full = cluster.load_cfg(None)
    cmd = cluster.job_command(full, ["--network", "ResNet18", "--approach", "cyclic"], node_rank=0, nnodes=1, nproc=8)
    assert "torch.distributed.run" in cmd and "--nproc-per-node=8" in cmd and "draco_b200.cli.distributed_nn" in cmd
    # local "cluster": check + idle work without ssh
    local = cluster.Cfg(dict(cluster.DEFAULT_CFG, nodes=["127.0.0.1"], state_file=str(tmp_path / "state.json")))
    assert cluster.check(local)["127.0.0.1"]["reachable"] in (True, False)
    assert cluster.idle(local) == {"127.0.0.1": True}


def test_bucket_plan_covers_arena_in_backprop_order():
    for name in ("LeNet", "ResNet18", "VGG11", "ResNet50"):
        L = ArenaLayout.from_model(build_model(name), bf16=True, channels_last=True)
        b = plan_buckets(L, 5)
        assert b[0][1] == L.ntiles and b[-1][0] == 0
        for (t0, t1, idxs), (u0, u1, jdxs) in zip(b[:-1], b[1:]):
            assert u1 == t0 and max(jdxs) < min(idxs)            # contiguous, later buckets hold earlier layers
        assert sorted(i for _, _, idxs in b for i in idxs) == list(range(L.ntensors))
        assert len(b) <= 8
        if name != "LeNet":
            assert (b[-1][1] - b[-1][0]) <= 0.12 * L.ntiles      # the exposed tail bucket is small


def test_ops_fall_back_to_aten_on_cpu():
    torch.manual_seed(0)
    x = torch.randn(4, 64, 8, 8)
    conv, ref = Conv2d(64, 128, 1, bias=False), torch.nn.Conv2d(64, 128, 1, bias=False)
    ref.load_state_dict(conv.state_dict())
    assert torch.equal(conv(x), ref(x))
    bn, rbn = FusedBatchNorm2d(64), torch.nn.BatchNorm2d(64)
    res = torch.randn_like(x)
    y = bn(x, residual=res, relu=True)
    assert torch.allclose(y, F.relu(rbn(x) + res), atol=1e-6)
    assert torch.allclose(bn.running_var, rbn.running_var) and int(bn.num_batches_tracked) == 1
    lin, rl = Linear(64, 10), torch.nn.Linear(64, 10)
    rl.load_state_dict(lin.state_dict())
    assert torch.equal(lin(x.mean((2, 3))), rl(x.mean((2, 3))))
    # state_dict keys are the reference's (nn.Module naming)
    keys = set(build_model("ResNet18").state_dict())
    assert {"conv1.weight", "bn1.running_mean", "layer1.0.conv1.weight", "layer2.0.shortcut.1.weight", "linear.bias"} <= keys
    vkeys = set(build_model("VGG11").state_dict())
    assert {"features.0.weight", "features.1.running_var", "features.4.weight", "classifier.1.weight", "classifier.6.bias"} <= vkeys


def test_fork_and_pool_fallbacks_on_cpu():
    """``Conv2d(x, fork=True)`` (residual blocks: the shortcut's gradient is folded into the dgrad epilogue on the GPU path) and
    ``global_avg_pool`` behave like the plain modules where the native kernels do not apply: same values, same gradients."""
    from draco_b200.models.resnet import BasicBlock
    from draco_b200.ops.pool import backend_counters, global_avg_pool
    torch.manual_seed(1)
    conv = Conv2d(8, 8, 3, padding=1, bias=False)
    x = torch.randn(2, 8, 6, 6, requires_grad=True)
    y, xf = conv(x, fork=True)
    (y.sum() + (xf * 2.0).sum()).backward()
    x2 = x.detach().clone().requires_grad_(True)
    (F.conv2d(x2, conv.weight, padding=1).sum() + (x2 * 2.0).sum()).backward()
    assert torch.allclose(x.grad, x2.grad, atol=1e-6)
    blk = BasicBlock(8, 8, 1)
    xb = torch.randn(2, 8, 6, 6, requires_grad=True)
    out = blk(xb)
    ref = F.relu(blk.bn2(blk.conv2(F.relu(blk.bn1(blk.conv1(xb))))) + xb)       # modules in training mode: same batch statistics
    assert torch.allclose(out, ref, atol=1e-5)
    before = backend_counters["aten"]
    z = torch.randn(3, 16, 4, 4)
    assert torch.allclose(global_avg_pool(z), F.avg_pool2d(z, 4).flatten(1), atol=1e-6)
    assert backend_counters["aten"] == before + 1


def test_profile_phases_reports_reference_timers(tmp_path, capsys):
    """--profile-phases: the reference's per-step timers (Comm / Comp / Encode on the worker, Method / Update on the PS)
    land in the JSONL records and in the human lines."""
    import json

    import torch

    from draco_b200 import JobConfig
    from draco_b200.parallel.trainer import Trainer
    mf = str(tmp_path / "m_{rank}.jsonl")
    cfg = JobConfig(network="LeNet", dataset="MNIST", approach="maj_vote", mode="maj_vote", group_size=3, num_workers=3,
                    worker_fail=1, err_mode="rev_grad", batch_size=16, max_steps=3, transport="gloo", synthetic_size=256,
                    eval_freq=10 ** 6, profile_phases=True, metrics_file=mf, log_interval=1, train_dir=str(tmp_path) + "/")
    t = Trainer(cfg, rank=0, world=1, device=torch.device("cpu"), quiet=False)
    t.fit(3)
    t.close()
    recs = [json.loads(l) for l in open(mf.replace("{rank}", "0"))]
    w = [r for r in recs if r["role"] == "worker"]
    p = [r for r in recs if r["role"] == "ps"]
    assert len(w) == 3 and len(p) == 3
    for k in ("t_fetch", "t_comp", "t_encode", "t_comm"):
        assert all(r[k] >= 0 for r in w), k
    assert all(r["t_comp"] > 0 for r in w)
    assert all(r["t_decode"] > 0 and r["t_update"] > 0 for r in p)
    out = capsys.readouterr().out
    assert "Comp:" in out and "Encode:" in out and "Method Time Cost" in out and "Update Time Cost" in out


def test_native_loader_staging_equals_python_staging(monkeypatch):
    """stage_batches through the C++ thread pool (gather + CIFAR augmentation) feeds the very same pixels as the Python path."""
    import torch

    from draco_b200 import JobConfig
    from draco_b200.parallel.trainer import Trainer

    def run(native):
        monkeypatch.setenv("DRACO_NATIVE_LOADER", "1" if native else "0")
        cfg = JobConfig(network="VGG11", dataset="Cifar10", approach="maj_vote", mode="maj_vote", group_size=3, num_workers=3,
                        worker_fail=0, err_mode="none", batch_size=4, max_steps=3, transport="gloo", synthetic_size=64,
                        eval_freq=10 ** 6, augment=True, lr=0.01)
        t = Trainer(cfg, rank=0, world=1, device=torch.device("cpu"), quiet=True)
        losses = [t.train_step()["loss"] for _ in range(3)]
        used = getattr(t.engine.worker, "_loader", None) is not None
        x = t.engine.worker.x_u8[1][0].clone()
        t.close()
        return losses, used, x

    l1, used1, x1 = run(True)
    l0, used0, x0 = run(False)
    assert used1 and not used0
    assert torch.equal(x1, x0) and l1 == l0


def test_load_dataset_reads_real_mnist_files_when_present(tmp_path):
    """`load_dataset` prefers on-disk torchvision data (never downloads): feed it a miniature MNIST in the raw idx format."""
    import struct

    import numpy as np
    import torch

    from draco_b200.data import load_dataset
    raw = tmp_path / "mnist_data" / "MNIST" / "raw"
    raw.mkdir(parents=True)
    rs = np.random.RandomState(0)

    def write(prefix, n):
        imgs = rs.randint(0, 256, size=(n, 28, 28), dtype=np.uint8)
        labs = rs.randint(0, 10, size=(n,), dtype=np.uint8)
        (raw / f"{prefix}-images-idx3-ubyte").write_bytes(struct.pack(">IIII", 2051, n, 28, 28) + imgs.tobytes())
        (raw / f"{prefix}-labels-idx1-ubyte").write_bytes(struct.pack(">II", 2049, n) + labs.tobytes())
        return imgs, labs

    tr_i, tr_l = write("train", 64)
    te_i, te_l = write("t10k", 16)
    ds = load_dataset("MNIST", root=str(tmp_path), train=True)
    assert not ds.synthetic and tuple(ds.images.shape) == (64, 1, 28, 28) and ds.images.dtype == torch.uint8
    assert np.array_equal(ds.images[:, 0].numpy(), tr_i) and np.array_equal(ds.labels.numpy(), tr_l)
    test = load_dataset("MNIST", root=str(tmp_path), train=False)
    assert len(test) == 16 and np.array_equal(test.labels.numpy(), te_l)
    x, y = ds.get_batch([3, 5, 7])
    assert x.shape == (3, 1, 28, 28) and y.tolist() == tr_l[[3, 5, 7]].tolist()
    # nothing on disk -> the synthetic stand-in of the same shape
    syn = load_dataset("MNIST", root=str(tmp_path / "nowhere"), synthetic_size=128)
    assert syn.synthetic and tuple(syn.images.shape) == (128, 1, 28, 28)


def test_cluster_run_idle_kill_by_exact_process_group(tmp_path, monkeypatch):
    """`cluster run` records the PID of every launched agent (its own session via setsid); `idle` polls exactly that PID and
    `kill` signals exactly that process group -- never a name pattern (reference: ps-aux based idle detection / pkill)."""
    import os
    import time

    from draco_b200.cli import cluster
    cfg = cluster.load_cfg(None)
    cfg.update({"nodes": ["localhost"], "remote_dir": str(tmp_path), "state_file": str(tmp_path / "state.json"), "gpus_per_node": 1})
    monkeypatch.setattr(cluster, "job_command",
                        lambda c, args, rank, nnodes, nproc: f"cd {c['remote_dir']} && {{ setsid nohup sleep 60 > job.log 2>&1 < /dev/null & echo $!; }}")
    pids = cluster.run(cfg, ["--max-steps", "1"])
    pid = pids["localhost"]
    assert pid > 0 and os.path.exists(cfg["state_file"])
    for _ in range(100):                                                       # setsid() happens right after the fork
        if os.getpgid(pid) == pid:
            break
        time.sleep(0.02)
    assert os.getpgid(pid) == pid and os.getpgid(pid) != os.getpgid(0)        # own group: killing it cannot touch the caller
    assert cluster.idle(cfg) == {"localhost": False}
    cluster.kill(cfg)
    for _ in range(50):
        if cluster.idle(cfg)["localhost"]:
            break
        time.sleep(0.1)
    assert cluster.idle(cfg) == {"localhost": True}
    monkeypatch.undo()
    text = cluster.job_command(cfg, ["--network", "LeNet"], 0, 1, 8)
    assert "setsid nohup" in text and "draco_b200.cli.distributed_nn --network LeNet" in text and text.rstrip().endswith("echo $!; }")


def test_replica_dropout_cpu_masks_are_keyed():
    """CPU fallback of ops.dropout.ReplicaDropout: same key -> same mask, different step / batch / layer -> different mask;
    without a context it is plain nn.Dropout (and identity in eval mode)."""
    import torch
    from draco_b200.ops import dropout as D
    l1, l2 = D.ReplicaDropout(0.5, salt=1), D.ReplicaDropout(0.5, salt=2)
    x = torch.ones(64, 32)
    D.set_context(5, 428, 2)
    a, b = l1(x), l1(x)
    assert torch.equal(a, b) and 0.3 < (a == 0).float().mean() < 0.7 and set(a.unique().tolist()) <= {0.0, 2.0}
    assert not torch.equal(l2(x), a)
    D.set_context(6, 428, 2)
    assert not torch.equal(l1(x), a)
    D.set_context(5, 428, 3)
    assert not torch.equal(l1(x), a)
    D.clear_context()
    l1.eval()
    assert torch.equal(l1(x), x)


class _StubEc2:
    """In-memory stand-in for a boto3 EC2 client: spot requests are fulfilled on the second poll, instances pass their status
    checks on the second poll -- enough to exercise the whole Ec2Fleet lifecycle without AWS."""

    def __init__(self):
        self.inst, self.reqs, self.polls, self.calls = {}, {}, {}, []

    def _new(self, n, name=None):
        ids = []
        for _ in range(n):
            iid = f"i-{len(self.inst):04d}"
            self.inst[iid] = {"InstanceId": iid, "State": {"Name": "pending"}, "PrivateIpAddress": f"10.0.0.{len(self.inst) + 1}",
                              "LaunchTime": len(self.inst), "Tags": [{"Key": "Name", "Value": name}] if name else []}
            ids.append(iid)
        return ids

    def run_instances(self, MinCount, MaxCount, TagSpecifications, **spec):
        self.calls.append(("run_instances", spec))
        ids = self._new(MaxCount, TagSpecifications[0]["Tags"][0]["Value"])
        return {"Instances": [{"InstanceId": i} for i in ids]}

    def request_spot_instances(self, SpotPrice, InstanceCount, Type, LaunchSpecification):
        self.calls.append(("request_spot_instances", SpotPrice, LaunchSpecification))
        out = []
        for _ in range(InstanceCount):
            rid = f"sir-{len(self.reqs):04d}"
            self.reqs[rid] = {"SpotInstanceRequestId": rid, "State": "open", "InstanceId": None}
            out.append(dict(self.reqs[rid]))
        return {"SpotInstanceRequests": out}

    def describe_spot_instance_requests(self, SpotInstanceRequestIds=None, Filters=None):
        ids = SpotInstanceRequestIds or list(self.reqs)
        for rid in ids:
            self.polls[rid] = self.polls.get(rid, 0) + 1
            if SpotInstanceRequestIds and self.polls[rid] >= 2 and not self.reqs[rid]["InstanceId"]:
                self.reqs[rid].update(State="active", InstanceId=self._new(1)[0])
        return {"SpotInstanceRequests": [dict(self.reqs[r]) for r in ids]}

    def create_tags(self, Resources, Tags):
        for i in Resources:
            self.inst[i]["Tags"] = Tags

    def describe_instance_status(self, InstanceIds, IncludeAllInstances):
        out = []
        for i in InstanceIds:
            self.polls[i] = self.polls.get(i, 0) + 1
            if self.polls[i] >= 2:
                self.inst[i]["State"] = {"Name": "running"}
            ok = "ok" if self.polls[i] >= 2 else "initializing"
            out.append({"InstanceId": i, "InstanceState": self.inst[i]["State"], "InstanceStatus": {"Status": ok}, "SystemStatus": {"Status": ok}})
        return {"InstanceStatuses": out}

    def describe_instances(self, Filters):
        name = Filters[0]["Values"][0]
        mine = [i for i in self.inst.values() if any(t["Value"] == name for t in i.get("Tags", []))]
        return {"Reservations": [{"Instances": mine}]}

    def cancel_spot_instance_requests(self, SpotInstanceRequestIds):
        for r in SpotInstanceRequestIds:
            self.reqs[r]["State"] = "cancelled"

    def terminate_instances(self, InstanceIds):
        for i in InstanceIds:
            self.inst[i]["State"] = {"Name": "terminated"}


def test_ec2_fleet_lifecycle_spot_and_on_demand(tmp_path):
    """launch (spot requests -> wait fulfilled -> tag -> wait running + status checks), summaries, live get_hosts (PS first),
    idempotent re-launch, shutdown (cancel requests + terminate), on-demand launch, and the failure paths (reference:
    tools/pytorch_ec2.py:128-257, 656-819)."""
    from draco_b200.cli import cluster
    cfg = cluster.Cfg(dict(cluster.DEFAULT_CFG, name="job7", n_instances=3, spot_price="12.5", image_id="ami-1", key_name="k",
                           security_group=["sg-1"], poll_s=0.0, launch_timeout_s=5))
    stub = _StubEc2()
    fleet = cluster.Ec2Fleet(cfg, client=stub)
    ids = fleet.launch()
    assert len(ids) == 3 and fleet.summarize() == {"running": ids}
    kind, price, spec = stub.calls[0]
    assert kind == "request_spot_instances" and price == "12.5" and spec["SecurityGroupIds"] == ["sg-1"] and spec["ImageId"] == "ami-1"
    files = cluster.get_hosts(cfg, str(tmp_path), fleet)
    assert cfg["nodes"] == ["10.0.0.1", "10.0.0.2", "10.0.0.3"] and files["hosts"].splitlines()[0] == "10.0.0.1\tnode0"
    assert fleet.launch() == ids and len(stub.inst) == 3                       # idempotent: no second fleet
    res = fleet.shutdown()
    assert sorted(res["terminated"]) == ids and fleet.summarize() == {"terminated": ids} and len(res["cancelled_requests"]) == 3
    od = cluster.Ec2Fleet(cluster.Cfg(dict(cfg, name="job8", spot_price="", n_instances=2)), client=stub)
    assert len(od.launch()) == 2 and stub.calls[-1][0] == "run_instances"
    # a request that fails terminally raises instead of polling forever; a fleet that never comes up times out
    bad = _StubEc2()
    bad.describe_spot_instance_requests = lambda **kw: {"SpotInstanceRequests": [
        {"SpotInstanceRequestId": r, "State": "failed", "InstanceId": None, "Status": {"Code": "price-too-low"}} for r in kw["SpotInstanceRequestIds"]]}
    with pytest.raises(RuntimeError, match="price-too-low"):
        cluster.Ec2Fleet(cfg, client=bad).launch()
    slow = _StubEc2()
    slow.describe_instance_status = lambda **kw: {"InstanceStatuses": []}
    with pytest.raises(TimeoutError):
        cluster.Ec2Fleet(cluster.Cfg(dict(cfg, spot_price="", launch_timeout_s=0.05)), client=slow).launch()


def test_cluster_parallel_fanout_and_single_job_default(tmp_path, monkeypatch):
    from draco_b200.cli import cluster
    cfg = cluster.Cfg(dict(cluster.DEFAULT_CFG, nodes=["127.0.0.1", "localhost"], state_file=str(tmp_path / "s.json"),
                           remote_dir=str(tmp_path)))
    out = cluster.run_parallel(cfg, "echo hello-$((1+1))")
    assert set(out) == {"127.0.0.1", "localhost"} and all(r["rc"] == 0 and "hello-2" in r["stdout"] for r in out.values())
    launched = []
    monkeypatch.setattr(cluster, "run_on", lambda c, n, cmd, t=60: launched.append((n, cmd)) or __import__("subprocess").CompletedProcess([], 0, "123\n", ""))
    assert list(cluster.run(cfg, ["--max-steps", "1"])) == ["127.0.0.1"]        # ONE job on the first node
    launched.clear()
    pids = cluster.run(cfg, ["--max-steps", "1"], all_nodes=True)
    assert list(pids) == ["127.0.0.1", "localhost"] and "job_node1.log" in launched[1][1] and "--nnodes=1" in launched[1][1]
    launched.clear()
    cluster.run(cfg, [], nnodes=2)
    assert "--node-rank=1" in launched[1][1] and "--nnodes=2" in launched[1][1]
