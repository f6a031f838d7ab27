"""CPU: the per-clip metric (measure.py:46-91) and the clip-sharded reduction, incl. a world_size-2 gloo run."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mumpy_oracle as orc


def measure_py_literal(result_binary, gt_mask):
    """measure.py:77-91 + :46-62 transcribed with numpy on boolean masks (the reference's own arithmetic)."""
    recall = np.sum(gt_mask & result_binary) / np.sum(gt_mask + 1e-6)
    precision = np.sum(gt_mask & result_binary) / (np.sum(result_binary) + 1e-6)
    f1 = 2 * (precision * recall) / (precision + recall + 1e-6)
    smooth = 1e-5
    iou = ((result_binary & gt_mask).sum() + smooth) / ((result_binary | gt_mask).sum() + smooth)
    return f1, iou


def _masks(n, seed):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand((n, 224, 224), generator=g) > 0.6
    gt = torch.rand((n, 224, 224), generator=g) > 0.7
    pred[0] = False          # empty prediction
    gt[1] = False            # empty ground truth
    return pred, gt


def test_metric_matches_measure_py():
    from mumpy_b200 import evaluate as ev
    pred, gt = _masks(6, 0)
    counts = orc.clip_counts(pred, gt)
    f1, iou = ev.f1_iou_from_counts(counts, 224 * 224)
    f1o, iouo = orc.f1_iou_from_counts(counts, 224 * 224)
    for i in range(6):
        a, b = measure_py_literal(pred[i].numpy(), gt[i].numpy())
        assert abs(float(f1[i]) - a) < 1e-12 and abs(float(iou[i]) - b) < 1e-12
        assert float(f1o[i]) == float(f1[i]) and float(iouo[i]) == float(iou[i])


def test_shard_bounds_cover_everything():
    from mumpy_b200 import evaluate as ev
    for n in (0, 1, 7, 1376):
        for world in (1, 2, 3, 8):
            spans = [ev.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mumpy_b200 import evaluate as ev
    pred, gt = _masks(n_clips, 3)
    logits_all = torch.where(pred, 1.0, -1.0).unsqueeze(1)

    def make_batch(lo, hi):
        return logits_all[lo:hi], gt[lo:hi]
    e = ev.ShardedEvaluator(predict=lambda x: x, make_batch=make_batch, n_clips=n_clips, micro_batch=3,
                            counts_fn=lambda lg, g: orc.clip_counts(lg[:, 0] > 0, g))
    r = e.run()
    table = ev.gather_count_table(r["counts"], n_clips)
    if rank == 0:
        q.put((r["f1"], r["iou"], r["n_valid"], table))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_reduction_equals_single_process():
    from mumpy_b200 import evaluate as ev
    n_clips = 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    f1, iou, n, table = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pred, gt = _masks(n_clips, 3)
    counts = orc.clip_counts(pred, gt)
    assert torch.equal(table, counts)
    s = ev.local_sums(counts, 224 * 224)
    assert n == int(s[2]) and abs(f1 - float(s[0] / s[2])) < 1e-12 and abs(iou - float(s[1] / s[2])) < 1e-12
