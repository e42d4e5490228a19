"""Cluster manager -- the role of the reference's ``tools/pytorch_ec2.py`` + shell glue (tools/local_script.sh,
remote_script.sh, update_git_dir.sh, hosts*), re-thought for GPU nodes.

The reference tool rents EC2 CPU spot instances with boto3, mounts EFS, writes ``hosts`` / ``hosts_alias`` /
``hosts_address`` and fans commands out with paramiko (tools/pytorch_ec2.py:93-971; its ``Cfg`` is a dict whose string
values are interpolated against the dict itself, :12-20).  What survives here:

  * ``Cfg``                 -- the same self-interpolating config dict
  * ``get_hosts``           -- write hosts / hosts_alias / hosts_address from a node list (or an EC2 query when boto3 exists)
  * ``check``               -- reachability + ``nvidia-smi`` inventory of every node
  * ``sync``                -- push the repo to every node (rsync over ssh; replaces remote_script.sh / update_git_dir.sh)
  * ``run``                 -- start one torchrun agent per node for a draco_b200 job (replaces mpirun --hostfile)
  * ``idle`` / ``kill``     -- find / stop running jobs (exact PIDs recorded at launch; never by pattern)
  * ``launch`` / ``shutdown`` -- EC2 instance lifecycle, only when boto3 is importable (it is not in this image)

One B200 node hosts the whole 1 PS + 7 workers job, so multi-node is mostly for sweeps: each node runs an independent
job unless ``--nnodes`` > 1 is given, in which case torchrun's rendezvous spans the nodes and the *collective*
transports are used across nodes (peer memory is intra-node).
"""
from __future__ import annotations

import argparse
import json
import os
import shlex
import subprocess
import sys
import time
from typing import Dict, List, Optional


class Cfg(dict):
    """dict whose string values may reference other keys as ``%(key)s`` (reference: tools/pytorch_ec2.py:12-20)."""

    def __getitem__(self, key):
        item = dict.__getitem__(self, key)
        if isinstance(item, str):
            for _ in range(8):
                new = item % self if "%(" in item else item
                if new == item:
                    break
                item = new
        return item


DEFAULT_CFG = Cfg({
    "name": "draco_b200",
    "nodes": ["127.0.0.1"],                 # hostnames / addresses of the GPU nodes
    "ssh_user": os.environ.get("USER", "root"),
    "ssh_key": "~/.ssh/id_rsa",
    "ssh_port": 22,
    "gpus_per_node": 8,
    "remote_dir": "/root/%(name)s",
    "python": sys.executable,
    "train_dir": "%(remote_dir)s/output/models/",
    "master_port": 29500,
    "state_file": ".cluster_state.json",
    # EC2 fields used only when boto3 is present
    "region": "us-west-2", "instance_type": "p6-b200.48xlarge", "image_id": "", "key_name": "", "n_instances": 1,
})


def load_cfg(path: Optional[str]) -> Cfg:
    cfg = Cfg(DEFAULT_CFG)
    if path:
        with open(path) as fh:
            cfg.update(json.load(fh))
    return cfg


def _is_local(host: str) -> bool:
    return host in ("127.0.0.1", "localhost")


def ssh_cmd(cfg: Cfg, host: str, command: str) -> List[str]:
    if _is_local(host):
        return ["bash", "-lc", command]
    return ["ssh", "-o", "StrictHostKeyChecking=no", "-o", "BatchMode=yes", "-p", str(cfg["ssh_port"]), "-i",
            os.path.expanduser(cfg["ssh_key"]), f"{cfg['ssh_user']}@{host}", command]


def run_on(cfg: Cfg, host: str, command: str, timeout: float = 60.0) -> subprocess.CompletedProcess:
    return subprocess.run(ssh_cmd(cfg, host, command), capture_output=True, text=True, timeout=timeout)


def get_hosts(cfg: Cfg, out_dir: str = ".") -> Dict[str, str]:
    """Write ``hosts`` (addr alias), ``hosts_alias`` and ``hosts_address`` like the reference's ``get_hosts``."""
    nodes = list(cfg["nodes"])
    files = {
        "hosts": "".join(f"{n}\tnode{i}\n" for i, n in enumerate(nodes)),
        "hosts_alias": "".join(f"node{i}\n" for i in range(len(nodes))),
        "hosts_address": "".join(f"{n}\n" for n in nodes),
    }
    for name, text in files.items():
        with open(os.path.join(out_dir, name), "w") as fh:
            fh.write(text)
    return files


def check(cfg: Cfg) -> Dict[str, dict]:
    out = {}
    for n in cfg["nodes"]:
        try:
            r = run_on(cfg, n, "nvidia-smi --query-gpu=name,memory.total --format=csv,noheader || echo NO_GPU", 30)
            gpus = [l for l in r.stdout.strip().splitlines() if l and "NO_GPU" not in l]
            out[n] = {"reachable": r.returncode == 0, "gpus": gpus}
        except (subprocess.TimeoutExpired, OSError) as e:
            out[n] = {"reachable": False, "error": str(e), "gpus": []}
    return out


def sync(cfg: Cfg, src: str = ".") -> None:
    for n in cfg["nodes"]:
        if _is_local(n) and os.path.abspath(src) == os.path.abspath(os.path.expanduser(cfg["remote_dir"])):
            continue
        dst = cfg["remote_dir"] if _is_local(n) else f"{cfg['ssh_user']}@{n}:{cfg['remote_dir']}"
        ssh = f"ssh -p {cfg['ssh_port']} -i {os.path.expanduser(cfg['ssh_key'])} -o StrictHostKeyChecking=no"
        cmd = ["rsync", "-az", "--delete", "--exclude", ".git", "--exclude", "gpurun_out", "--exclude", "build", "-e", ssh,
               src.rstrip("/") + "/", dst]
        if _is_local(n):
            cmd = ["rsync", "-a", "--exclude", ".git", src.rstrip("/") + "/", dst]
        subprocess.run(cmd, check=True)


def job_command(cfg: Cfg, job_args: List[str], node_rank: int, nnodes: int, nproc: int) -> str:
    master = cfg["nodes"][0]
    tr = (f"{cfg['python']} -m torch.distributed.run --nnodes={nnodes} --node-rank={node_rank} --nproc-per-node={nproc} "
          f"--master-addr {master if nnodes > 1 else '127.0.0.1'} --master-port {cfg['master_port']} "
          f"-m draco_b200.cli.distributed_nn {' '.join(shlex.quote(a) for a in job_args)}")
    log = f"{cfg['remote_dir']}/job_node{node_rank}.log"
    # setsid: the agent leads its own session / process group, so ``kill`` can address exactly that group by the recorded PID
    # (the braces matter: only the agent is backgrounded, with all three standard streams redirected, so the launching shell --
    # and the ssh / subprocess pipe behind it -- returns immediately with the PID instead of waiting for the job)
    return f"cd {cfg['remote_dir']} && {{ setsid nohup {tr} > {log} 2>&1 < /dev/null & echo $!; }}"


def run(cfg: Cfg, job_args: List[str], nnodes: int = 1, nproc: Optional[int] = None) -> Dict[str, int]:
    """Start the job; returns {node: pid of the launched agent} and records it for ``kill``."""
    nproc = nproc or int(cfg["gpus_per_node"])
    pids = {}
    nodes = cfg["nodes"][:nnodes] if nnodes > 1 else cfg["nodes"]
    for i, n in enumerate(nodes):
        r = run_on(cfg, n, job_command(cfg, job_args, i if nnodes > 1 else 0, nnodes, nproc), 60)
        pids[n] = int(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else -1
    with open(cfg["state_file"], "w") as fh:
        json.dump({"pids": pids, "started": time.time(), "args": job_args}, fh)
    return pids


def idle(cfg: Cfg) -> Dict[str, bool]:
    """A node is idle when none of the recorded agent PIDs is alive (reference: idle detection via ``ps aux``)."""
    state = _state(cfg)
    res = {}
    for n in cfg["nodes"]:
        pid = state.get("pids", {}).get(n, -1)
        if pid <= 0:
            res[n] = True
            continue
        r = run_on(cfg, n, f"kill -0 {pid} 2>/dev/null && echo BUSY || echo IDLE", 30)
        res[n] = "IDLE" in r.stdout
    return res


def kill(cfg: Cfg) -> None:
    """Stop the recorded agents by exact PID / process group -- never by name pattern."""
    state = _state(cfg)
    for n, pid in state.get("pids", {}).items():
        if pid > 0:
            # the recorded PID is a process-group leader (launched under setsid): signal that group, and only that group
            run_on(cfg, n, f"kill -TERM -- -{pid} 2>/dev/null || kill -TERM {pid} 2>/dev/null; true", 30)


def _state(cfg: Cfg) -> dict:
    try:
        with open(cfg["state_file"]) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return {}


def _ec2(cfg: Cfg, action: str) -> None:
    try:
        import boto3  # noqa: F401
    except ImportError:
        raise SystemExit("boto3 is not installed in this image: EC2 launch/shutdown are unavailable; "
                         "list the nodes in the config instead (\"nodes\": [...])")
    ec2 = boto3.client("ec2", region_name=cfg["region"])
    if action == "launch":
        r = ec2.run_instances(ImageId=cfg["image_id"], InstanceType=cfg["instance_type"], KeyName=cfg["key_name"],
                              MinCount=int(cfg["n_instances"]), MaxCount=int(cfg["n_instances"]),
                              TagSpecifications=[{"ResourceType": "instance", "Tags": [{"Key": "Name", "Value": cfg["name"]}]}])
        print(json.dumps([i["InstanceId"] for i in r["Instances"]]))
    else:
        r = ec2.describe_instances(Filters=[{"Name": "tag:Name", "Values": [cfg["name"]]}])
        ids = [i["InstanceId"] for res in r["Reservations"] for i in res["Instances"]]
        if ids:
            ec2.terminate_instances(InstanceIds=ids)
        print(json.dumps(ids))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="draco_b200 cluster manager")
    ap.add_argument("command", choices=["get_hosts", "check", "sync", "run", "idle", "kill", "launch", "shutdown", "show_cfg"])
    ap.add_argument("--config", default=None, help="JSON file overriding the default Cfg")
    ap.add_argument("--nnodes", type=int, default=1)
    ap.add_argument("--nproc-per-node", type=int, default=None)
    ap.add_argument("job_args", nargs=argparse.REMAINDER, help="flags forwarded to draco_b200.cli.distributed_nn (after --)")
    a = ap.parse_args(argv)
    cfg = load_cfg(a.config)
    job_args = [x for x in a.job_args if x != "--"]
    if a.command == "show_cfg":
        print(json.dumps({k: cfg[k] for k in cfg}, indent=1))
    elif a.command == "get_hosts":
        print(json.dumps(get_hosts(cfg)))
    elif a.command == "check":
        print(json.dumps(check(cfg), indent=1))
    elif a.command == "sync":
        sync(cfg)
    elif a.command == "run":
        print(json.dumps(run(cfg, job_args, a.nnodes, a.nproc_per_node)))
    elif a.command == "idle":
        print(json.dumps(idle(cfg)))
    elif a.command == "kill":
        kill(cfg)
    else:
        _ec2(cfg, a.command)
    return 0


if __name__ == "__main__":
    sys.exit(main())
