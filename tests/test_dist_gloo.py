"""Multi-rank path on the CPU (gloo, world_size 2): tile sharding, the counter
all-reduce and the per-lane reports.  The per-tile counters each rank
contributes come from the C oracle, so the printed result can be compared with
the reference's stdout for the same run."""
import io
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import GOLDEN, load_manifest, parse_count_args

MAN = load_manifest()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tile_rows(case, o, lanes, tiles, mine):
    from oracle import c_port as CP
    from oracle import ref_port as R
    targets = R.parse_target_file(os.path.join(GOLDEN, case["targets"]), levels=o["levels"] + 1, limit=o["limit"])
    centres = [t[0][0] for t in targets]
    offs, idx = [0], []
    for t in targets:
        for ring in t[1:]:
            idx.extend(ring)
            offs.append(len(idx))
    rows = []
    for ordinal in mine:
        lane, tile = lanes[ordinal // len(tiles)], tiles[ordinal % len(tiles)]
        planes, kinds = [], []
        for s, e in o["ranges"]:
            p, k, filt, n = R.load_tile_planes(os.path.join(GOLDEN, case["run"]), lane, tile, s, e)
            planes += p
            kinds += k
        rows.append(CP.count_tile(planes, kinds, filt, centres, offs, idx, o["levels"], o["edit"], o["hamming"],
                                  want_per_target=False)[1])
    return np.array(rows, dtype=np.int64).reshape(len(mine), 1 + 5 * o["levels"]), len(targets)


def _worker(rank, world, port, case, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_port as R
    from well_duplicates_b200 import flowcell
    o = parse_count_args(case["args"])
    lanes = o["lanes"].split(",")
    tiles = R.tile_list(o["stype"], o["tiles"])
    mine = flowcell.plan(len(lanes), len(tiles), rank, world)
    rows, n_targets = _tile_rows(case, o, lanes, tiles, mine)
    buf = flowcell.exchange_host(rows, mine, len(lanes), len(tiles), 1 + 5 * o["levels"], dist)
    text = io.StringIO()
    flowcell.print_reports(text, lanes, tiles, buf, n_targets, o["levels"], verbose=not o["summary"])
    with open("%s.%d" % (out_path, rank), "w") as fh:
        fh.write(text.getvalue())
    np.save("%s.%d.npy" % (out_path, rank), buf)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["two_lanes", "lev_default", "summary", "regex_tiles"])
def test_two_ranks_reproduce_reference_report(name, tmp_path):
    case = [c for c in MAN["count"] if c["name"] == name][0]
    out = str(tmp_path / "report")
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    with open(os.path.join(GOLDEN, "count", name + ".stdout")) as fh:
        want = fh.read()
    for rank in (0, 1):
        with open("%s.%d" % (out, rank)) as fh:
            assert fh.read() == want                     # every rank holds the full result
    a, b = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.array_equal(a, b)


def test_plan_covers_every_tile_once():
    from well_duplicates_b200 import flowcell
    for lanes, tpl, world in ((8, 96, 8), (8, 96, 3), (1, 5, 4), (2, 1, 2), (4, 704, 8)):
        seen = np.concatenate([flowcell.plan(lanes, tpl, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(lanes * tpl))
        sizes = [len(flowcell.plan(lanes, tpl, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_lane_rows_are_the_sum_of_their_tiles():
    from well_duplicates_b200 import flowcell
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 1000, size=(6, 11))
    buf = flowcell.exchange_host(rows, np.arange(6), 2, 3, 11, None)
    assert np.array_equal(buf[:6], rows)
    assert np.array_equal(buf[6], rows[:3].sum(axis=0)) and np.array_equal(buf[7], rows[3:].sum(axis=0))


def _driver_worker(rank, world, port, case, out_path):
    """flowcell.count_rank_tiles (tile walk + staging pipeline) on each rank, with the oracle-backed
    stand-in for the engine, then the all-reduce and the report as the driver does them."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_cli_host_logic import OracleEngine
    from well_duplicates_b200 import count_cli, flowcell, reader, staging
    from well_duplicates_b200.targets import load_targets
    o = parse_count_args(case["args"])
    lanes = o["lanes"].split(",")
    tiles = count_cli.expected_tiles(o["stype"], o["tiles"])
    wanted = [c for s, e in o["ranges"] for c in range(s, e)]
    targets = load_targets(os.path.join(GOLDEN, case["targets"]), levels=o["levels"] + 1, limit=o["limit"])
    eng = OracleEngine()
    eng.load_targets(*targets.to_csr(o["levels"]), o["levels"])
    rd = reader.BCLReader(os.path.join(GOLDEN, case["run"]), engine=eng)
    st = staging.Stager(pinned=False, threads=2, cbcl_cache=rd._cbcl_cache)
    mine = flowcell.plan(len(lanes), len(tiles), rank, world)
    rows = flowcell.count_rank_tiles(eng, rd, st, lanes, tiles, mine, wanted, o["levels"], o["edit"], o["hamming"])
    st.close()
    buf = flowcell.exchange_host(rows, mine, len(lanes), len(tiles), 1 + 5 * o["levels"], dist)
    text = io.StringIO()
    flowcell.print_reports(text, lanes, tiles, buf, len(targets), o["levels"], verbose=not o["summary"])
    with open("%s.%d" % (out_path, rank), "w") as fh:
        fh.write(text.getvalue())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["two_lanes", "lev_default", "cbcl_default"])
def test_two_rank_driver_loop_reproduces_reference_report(name, tmp_path):
    case = [c for c in MAN["count"] if c["name"] == name][0]
    out = str(tmp_path / "report")
    mp.spawn(_driver_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    with open(os.path.join(GOLDEN, "count", name + ".stdout")) as fh:
        want = fh.read()
    for rank in (0, 1):
        with open("%s.%d" % (out, rank)) as fh:
            assert fh.read() == want
