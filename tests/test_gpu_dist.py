"""Two ranks, two GPUs, NCCL: the CUDA library behind the multi-GPU dispatcher.

Both ways of closing a replicate are run -- sums through the host (gloo side group) and the
no-round-trip path (per-chromosome sums all-reduced by NCCL on the library's own stream,
lambda / scale factor computed on the device) -- and both must reproduce the single-process
oracle bit for bit.  Needs two devices; skipped otherwise (the driver's 1-GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import util
from cases import BY_NAME
from genrich_b200 import capi
from genrich_b200.dist import ShardedEngine

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_name, out_dir, want_stats):
    import faulthandler
    faulthandler.enable()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    hg = td.new_group(backend="gloo")
    case = BY_NAME[case_name]
    api = capi.load_cuda()
    eng = ShardedEngine(api, case.chrom_len, util.case_params(case), torch.device("cuda", rank), host_group=hg,
                        exclusions=case.bed)
    for expt, ctrl, save in util.case_inputs(case):
        e, c = eng.route(expt), (None if ctrl is None else eng.route(ctrl))
        eng.replicate(lambda cx: cx.push_intervals(e),
                      None if c is None else (lambda cx: cx.push_intervals(c)), save, want_stats=want_stats)
    peaks, rs = eng.call_peaks()
    if rank == 0:
        np.save(os.path.join(out_dir, "peaks.npy"), peaks)
    eng.close()                              # torch's views of the library's stream and buffers go before the context
    td.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("want_stats", [True, False])
@pytest.mark.parametrize("name", ["c2_ctrl_q", "c5_multimap_ctrl_p", "fisher_missing_chrom", "bed_ctrl_q"])
def test_two_gpus_equal_oracle(name, want_stats, tmp_path):
    case = BY_NAME[name]
    _, ref, _ = util.run_case(util.oracle_api(), case)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path), want_stats), nprocs=2, join=True)
    peaks = np.load(os.path.join(tmp_path, "peaks.npy"))
    assert len(peaks) == len(ref.peaks) and len(peaks) > 0
    for f in ("chrom", "start", "end", "summit"):
        assert np.array_equal(peaks[f], ref.peaks[f]), f
    for f in ("pval", "qval"):
        assert np.allclose(peaks[f], ref.peaks[f], rtol=0, atol=1e-4), f
    # the AUC sums len * (-log10 q - threshold) over a peak: the 1e-4 bar of its terms, scaled
    assert np.allclose(peaks["auc"], ref.peaks["auc"], rtol=1e-4, atol=1e-3)
