"""Host-side sharding logic, including a world_size-2 gloo run on CPU."""
import json
import os
import subprocess
import sys
import textwrap

from multi_spatialmllm_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_is_balanced_and_deterministic():
    costs = {f"scene{i:04d}_00": shard.scene_cost(150_000 + 997 * (i * 37 % 101), 300 + (i * 13) % 90) for i in range(64)}
    a = shard.partition_lpt(costs, 8)
    b = shard.partition_lpt(dict(reversed(list(costs.items()))), 8)
    assert a == b
    assert sorted(s for part in a for s in part) == sorted(costs)
    loads = [sum(costs[s] for s in part) for part in a]
    assert max(loads) / (sum(loads) / 8) < 1.05
    assert shard.my_shard(costs, 3, 8) == a[3]
    assert shard.part_path("/x/out.jsonl", 7) == "/x/out.part-00007.jsonl"


def test_scene_cost_matches_survey_formula():
    assert shard.scene_cost(200_000, 330) == 200_000 * 330 + (330 * 329 // 2) * 200_000 / 8
    assert shard.scene_cost(1000, 10, n_pairs=5) == 1000 * 10 + 5 * 1000 / 8


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import json, os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from multi_spatialmllm_b200 import shard
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        costs = {{f"scene{{i:04d}}_00": shard.scene_cost(1000 * (i + 1), 50) for i in range(10)}}
        mine = shard.my_shard(costs, rank, world)
        part = shard.part_path({str(tmp_path / 'samples.jsonl')!r}, rank)
        with open(part, "w") as f:
            for s in mine:
                f.write(json.dumps({{"id": s, "rank": rank}}) + "\\n")
        merged = shard.gather_manifest({{"rank": rank, "part": part, "scenes": mine, "samples": len(mine)}},
                                       {str(tmp_path / 'manifest.json')!r})
        if rank == 0:
            assert len(merged) == world
            n = shard.merge_jsonl_parts([m["part"] for m in merged], {str(tmp_path / 'samples.jsonl')!r})
            assert n == 10
        else:
            assert merged is None
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    import socket
    with socket.socket() as sock:                       # a free port, so parallel test runs do not collide
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    man = json.load(open(tmp_path / "manifest.json"))
    assert man["total_samples"] == 10 and len(man["parts"]) == 2
    ids = [json.loads(l)["id"] for l in open(tmp_path / "samples.jsonl")]
    assert sorted(ids) == [f"scene{i:04d}_00" for i in range(10)]
