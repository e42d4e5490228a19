"""End-to-end runs of the command-line entry points on CPU (subprocesses, like a user would launch them)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="2")


def _run(args, timeout=420, cwd=None):
    r = subprocess.run([sys.executable, *args], capture_output=True, text=True, timeout=timeout, env=ENV, cwd=cwd or ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    return r.stdout + r.stderr


def test_distributed_nn_single_process_then_evaluator(tmp_path):
    """One process hosting the PS and every worker (the --no-cuda path), checkpoints every 2 steps, then the evaluator follows
    the run through its checkpoints (reference: run_pytorch.sh + evaluate_pytorch.sh)."""
    d = str(tmp_path) + "/"
    out = _run(["-m", "draco_b200.cli.distributed_nn", "--no-cuda", "--network", "FC", "--dataset", "MNIST", "--approach", "cyclic",
                "--worker-fail", "2", "--num-workers", "7", "--err-mode", "constant", "--batch-size", "4", "--lr", "0.01",
                "--momentum", "0.9", "--max-steps", "4", "--eval-freq", "2", "--train-dir", d, "--compress-grad", "compress",
                "--synthetic-size", "256", "--log-interval", "1"])
    assert "I am master" in out and "Worker:" in out and "Master Step" in out
    assert os.path.exists(d + "model_step_2") and os.path.exists(d + "model_step_4")
    ev = _run(["-m", "draco_b200.cli.distributed_evaluator", "--network", "FC", "--dataset", "MNIST", "--model-dir", d, "--eval-freq", "2",
               "--eval-batch-size", "64", "--device", "cpu", "--poll-s", "0.05", "--max-evals", "2", "--timeout-s", "20"])
    assert ev.count("Prec@1") >= 2, ev[-1500:]


def test_distributed_nn_launch_spawns_one_process_per_role(tmp_path):
    """--launch 3: 1 PS + 2 workers as three Gloo processes (the reference's `mpirun -n 3`), resumed from a checkpoint once."""
    d = str(tmp_path) + "/"
    common = ["-m", "draco_b200.cli.distributed_nn", "--launch", "3", "--master-port", "29641", "--no-cuda", "--network", "LeNet",
              "--dataset", "MNIST", "--approach", "maj_vote", "--mode", "maj_vote", "--group-size", "2", "--worker-fail", "0",
              "--err-mode", "rev_grad", "--batch-size", "8", "--eval-freq", "3", "--train-dir", d, "--synthetic-size", "128",
              "--compress-grad", "None", "--log-interval", "1"]
    out = _run(common + ["--max-steps", "3"])
    assert out.count("I am worker") == 2 and out.count("I am master") == 1
    assert os.path.exists(d + "model_step_3")
    out2 = _run(common + ["--max-steps", "5", "--checkpoint-step", "3"])
    assert "Step: 4" in out2 and "Step: 5" in out2 and "Step: 2," not in out2            # continued after the checkpoint


def test_single_machine_trainer_cli():
    out = _run(["-m", "draco_b200.cli.single_machine", "--network", "LeNet", "--dataset", "MNIST", "--max-steps", "3", "--batch-size",
                "16"])
    assert "Prec@1" in out or "loss" in out.lower()


def test_data_prepare_and_cluster_show_cfg(tmp_path):
    out = _run(["-m", "draco_b200.data.prepare", "--root", str(tmp_path), "--synthetic-size", "64"])
    assert os.path.exists(tmp_path / "synthetic_MNIST.pt") and os.path.exists(tmp_path / "synthetic_Cifar10.pt"), out
    cfg = _run(["-m", "draco_b200.cli.cluster", "show_cfg"])
    assert "nodes" in cfg and "remote_dir" in cfg
    hosts = _run(["-m", "draco_b200.cli.cluster", "get_hosts"], cwd=str(tmp_path))
    assert os.path.exists(tmp_path / "hosts_address"), hosts


def test_reference_example_job_scripts(tmp_path):
    """tools/run_pytorch.sh is the reference's shipped example (src/run_pytorch.sh: FC/MNIST, cyclic code, n=7, s=2, constant
    adversary, compression on); run it as 2 Gloo processes on CPU, then tools/evaluate_pytorch.sh on its checkpoint."""
    env = dict(ENV, NPROC="2", MAX_STEPS="2", PORT="29671")
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "run_pytorch.sh"), "--no-cuda", "--synthetic-size", "128", "--eval-freq", "2"],
                       capture_output=True, text=True, timeout=420, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "I am master" in r.stdout and os.path.exists(tmp_path / "output" / "models" / "model_step_2")
    env = dict(ENV, EVAL_FREQ="2")
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "evaluate_pytorch.sh"), "--device", "cpu", "--poll-s", "0.05", "--max-evals", "1",
                        "--timeout-s", "20", "--eval-batch-size", "64"], capture_output=True, text=True, timeout=200, env=env,
                       cwd=str(tmp_path))
    assert r.returncode == 0 and "Prec@1" in r.stdout, (r.stdout[-1500:], r.stderr[-2000:])


@pytest.mark.parametrize("name,flags", [
    ("mean+codec", ["--approach", "baseline", "--mode", "normal", "--worker-fail", "0", "--compress-grad", "compress"]),
    ("geomedian", ["--approach", "baseline", "--mode", "geometric_median", "--worker-fail", "1", "--err-mode", "rev_grad"]),
    ("krum+async", ["--approach", "baseline", "--mode", "krum", "--worker-fail", "1", "--err-mode", "constant", "--comm-type", "Async"]),
    ("cyclic", ["--approach", "cyclic", "--worker-fail", "1", "--err-mode", "random"]),
    ("vote+omniscient", ["--approach", "maj_vote", "--mode", "maj_vote", "--group-size", "3", "--worker-fail", "1", "--err-mode",
                         "omniscient"]),
])
def test_every_aggregation_rule_as_a_multi_process_job(tmp_path, name, flags):
    """1 PS + 5 workers packed onto 2 Gloo processes, 3 steps, for every aggregation rule / wire option."""
    port = 29700 + ["mean+codec", "geomedian", "krum+async", "cyclic", "vote+omniscient"].index(name)
    out = _run(["-m", "draco_b200.cli.distributed_nn", "--launch", "2", "--master-port", str(port), "--no-cuda", "--network", "LeNet",
                "--dataset", "MNIST", "--num-workers", "5", "--batch-size", "8", "--max-steps", "3", "--eval-freq", "1000",
                "--train-dir", str(tmp_path) + "/", "--synthetic-size", "128", "--log-interval", "1", "--compress-grad", "None", *flags])
    assert out.count("done at step 3") == 2, out[-2500:]
