"""The reference arm of bench.py (`--impl reference`: the reference's CPU path = oracle port on the host cores) runs without
a GPU and prints one JSON line with the keys the driver reads; host-side partition helpers of the C4 / C5 workloads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "equirect output Mpix/s" and d["unit"] == "Mpix/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("C2 ")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_stereo_assignment_and_shares_cover_the_frame_once():
    import octvr_b200 as vr

    class T:
        out_size = (256, 128)                                   # one eye: 256 x 128
    for world in (1, 2, 4, 8):
        asg = vr.sharding.stereo_assignment(world)
        assert len(asg) == world and sum(len(a) for a in asg) == max(world, 2)
        # rows of the packed top-bottom frame (2 eyes: 256 luma rows + 128 chroma rows) owned by the ranks: a partition
        rows = []
        for jobs in asg:
            for eye, b, per in jobs:
                y0, y1 = vr.sharding.row_bands(128, per, 32)[b]
                rows += list(range(eye * 128 + y0, eye * 128 + y1))
        assert sorted(rows) == list(range(256))
    assert vr.sharding.streams_of_rank(16, 3, 8) == [3, 11]
