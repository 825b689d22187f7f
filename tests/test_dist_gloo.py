"""world_size-2 test of the multi-GPU host logic on CPU (gloo).

The dispatcher in genrich_b200/dist.py is engine-agnostic: here it drives the CPU
oracle twin, so the chromosome sharding, the per-chromosome sum exchange, the
histogram all-gather and the peak gather are exercised without a GPU.  The result
must equal the single-process run bit for bit (lambda, factor, peaks, q)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import util
from cases import BY_NAME
from genrich_b200 import capi, host
from genrich_b200.dist import ShardedEngine


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    case = BY_NAME[case_name]
    api = util.oracle_api()
    eng = ShardedEngine(api, case.chrom_len, util.case_params(case), torch.device("cpu"), exclusions=case.bed)
    for expt, ctrl, save in util.case_inputs(case):
        e, c = eng.route(expt), (None if ctrl is None else eng.route(ctrl))
        eng.replicate(lambda cx: cx.push_intervals(e),
                      None if c is None else (lambda cx: cx.push_intervals(c)), save)
    peaks, rs = eng.call_peaks()
    if rank == 0:
        np.save(os.path.join(out_dir, "peaks.npy"), peaks)
        np.save(os.path.join(out_dir, "scal.npy"),
                np.array([[s.lambda_, s.factor] for s in eng.sample_stats], dtype=np.float32))
    owned = [c for c in range(len(case.chrom_len)) if eng.owned[c]]
    assert len(owned) < len(case.chrom_len)
    assert len(owned) > 0 or world > len(case.chrom_len)
    td.destroy_process_group()


@pytest.mark.parametrize("name", ["c2_ctrl_q", "c4_fisher_q", "fisher_missing_chrom", "c5_multimap_ctrl_p", "bed_fisher_q"])
def test_two_ranks_equal_one(name, tmp_path):
    case = BY_NAME[name]
    _, ref, _ = util.run_case(util.oracle_api(), case)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    peaks = np.load(os.path.join(tmp_path, "peaks.npy"))
    scal = np.load(os.path.join(tmp_path, "scal.npy"))
    assert peaks.tobytes() == ref.peaks.tobytes()
    want = np.array([[s.lambda_, s.factor] for s in ref.sample_stats], dtype=np.float32)
    assert np.array_equal(scal.view(np.uint32), want.view(np.uint32))


def test_more_ranks_than_chromosomes(tmp_path):
    """Four ranks, three chromosomes: one rank owns nothing and still takes part in every exchange
    (sums, histogram all-gather with an empty list, peak gather with an empty list)."""
    for name in ("c2_ctrl_q", "bed_fisher_q"):
        case = BY_NAME[name]
        _, ref, _ = util.run_case(util.oracle_api(), case)
        port = _free_port()
        mp.spawn(_worker, args=(4, port, name, str(tmp_path)), nprocs=4, join=True)
        peaks = np.load(os.path.join(tmp_path, "peaks.npy"))
        assert peaks.tobytes() == ref.peaks.tobytes() and len(peaks) > 0


def test_lpt_shard_balances():
    L = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
         138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
         83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
    for w in (2, 4, 8):
        own = host.lpt_shard(L, w)
        load = np.bincount(own, weights=np.asarray(L, dtype=np.float64), minlength=w)
        assert load.max() / load.mean() < 1.04
