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


def test_merge_parquet_parts_restores_scene_order(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    from multi_spatialmllm_b200 import tables
    order = [f"scene{i:04d}_00" for i in range(6)]
    parts = [str(tmp_path / f"t.part-{r:05d}.parquet") for r in range(3)]
    for r, mine in enumerate((order[4::-3], order[2:4], [])):          # rank 0: scenes 4, 1; rank 1: 2, 3; rank 2: none
        if not mine:
            continue
        w = None
        for s in mine:
            t = pa.table({"key": [f"{s}:image_to_points:{k:05d}" for k in range(3)], "values": ["[]"] * 3})
            w = w or pq.ParquetWriter(parts[r], t.schema)
            w.write_table(t)
        w.close()
    out = str(tmp_path / "t.parquet")
    assert tables.merge_parquet_parts(parts, out, order) == 12
    keys = pq.read_table(out)["key"].to_pylist()
    assert [k.split(":")[0] for k in keys[::3]] == [order[1], order[2], order[3], order[4]]


def test_two_rank_tables_cli_gloo(tmp_path):
    """`python -m multi_spatialmllm_b200.tables relations` under torchrun with 2 ranks: every rank writes its part, rank 0
    merges them in the split's scene order (the per-scene geometry is stubbed — this box has no GPU; the kernels' side is
    covered by tests/test_handler_tables.py on the GPU)."""
    import pickle
    import numpy as np
    import pandas as pd
    inst = tmp_path / "data" / "scannet" / "scannet_instance_data"
    infos = {}
    for i in range(5):
        sid = f"scene{i:04d}_00"
        (inst / sid).mkdir(parents=True)
        np.save(inst / sid / "aligned_points.npy", np.zeros((100 * (i + 1), 6)))
        infos[sid] = {"num_posed_images": 3 + i, "intrinsic_matrix": np.eye(4), "axis_align_matrix": np.eye(4), "num_objects": 0,
                      "images_info": {f"{5 * k:05d}": {"extrinsic_matrix": np.eye(4)} for k in range(3 + i)}}
    info = inst / "scenes_val_info_i_D5.pkl"
    with open(info, "wb") as f:
        pickle.dump(infos, f)
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import pyarrow as pa
        from multi_spatialmllm_b200 import tables

        def fake_relations(h, output_parquet, scene_ids=None, device="cuda:0", warning_file=None):
            full = tables._ParquetStream(output_parquet)
            nz = tables._ParquetStream(output_parquet.replace(".parquet", "_nonzero.parquet"))
            for s in scene_ids:
                n = h.get_num_posed_images(s)
                t = pa.table({{"scene_id": [s] * n, "image_id1": ["00000"] * n, "image_id2": ["00005"] * n,
                              "overlap": [float(k) for k in range(n)], "distance": [0.0] * n, "yaw": [0.0] * n, "pitch": [0.0] * n}})
                full.write(t)
                nz.write(t.slice(1))
            return {{"rows": full.close(), "rows_nonzero": nz.close()}}

        tables.run_relations_split = fake_relations
        res = tables.main(["relations", "--splits", "val", "--val_info_path", {str(info)!r}, "--val_output", {str(tmp_path / 'cam.parquet')!r},
                           "--instance_data_root", {str(inst)!r}])
        if int(os.environ["RANK"]) == 0:
            assert res["val"]["rows"] == sum(range(3, 8)) and res["val"]["rows_nonzero"] == sum(range(2, 7)), res
    """))
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert r.returncode == 0, r.stderr[-3000:]
    df = pd.read_parquet(tmp_path / "cam.parquet")
    assert list(dict.fromkeys(df["scene_id"])) == [f"scene{i:04d}_00" for i in range(5)]       # the split's scene order
    assert len(df) == sum(range(3, 8)) and len(pd.read_parquet(tmp_path / "cam_nonzero.parquet")) == sum(range(2, 7))
    assert os.path.exists(tmp_path / "cam.part-00000.parquet") and os.path.exists(tmp_path / "cam.part-00001.parquet")
