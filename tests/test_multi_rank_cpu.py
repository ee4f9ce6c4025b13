"""world_size-2 gloo test (CPU) of the N>1 host logic of bench.py: every rank gets its own shard of
utterances (weak scaling, no data-path collective) and the job-level numbers are reduced as
bench.py reduces them (max of times, sum of work)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_shard_and_reduce(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        import bench
        from oracle import vits as ov
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        hp = ov.tiny_hparams()
        utts, raw = bench.make_batch(hp, 4, seed=100 + rank)
        # shards differ between ranks (different seeds) but have the same batch size
        tx = torch.tensor([u["x_tst"].size for u in utts], dtype=torch.int64)
        gathered = [torch.zeros_like(tx) for _ in range(world)]
        dist.all_gather(gathered, tx)
        t = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)
        tot = torch.tensor([float(tx.sum())], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(json.dumps({{"shards_differ": not torch.equal(gathered[0], gathered[1]), "tmax": t.tolist(),
                              "total": float(tot), "sum_check": float(sum(int(g.sum()) for g in gathered))}}))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["shards_differ"] and d["tmax"] == [11.0, 20.0] and d["total"] == d["sum_check"]


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--tiny"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["metric"] == "audio-sec/sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
