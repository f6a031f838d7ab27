"""Clip-sharded evaluation: the thin replacement of the reference's test.py hot loop (test.py:86-111) and of
measure.py's per-image metric (measure.py:46-91,121-130), one process per GPU.

Inference shards by clip (test.py runs batch_size=1, clips are independent): every rank owns a contiguous slice of
the clip list, runs the forward on its own GPU with no data-path collective, reduces logits to integer counts
[TP, n_pred, n_gt, n_union] on the device (mumpy_mask_counts), turns them into per-clip F1 / IoU in float64 exactly as
measure.py does, and a single all-reduce (3 x fp64 = 24 bytes; NCCL over NVLink on the GPU box, gloo in CPU tests)
produces the split means.  The per-clip count table can also be all-gathered for an order-independent audit.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def f1_iou_from_counts(counts: torch.Tensor, n_pixels: int):
    """measure.py:77-91 + iou_score :46-62 from integer counts (N,4) = [TP, n_pred, n_gt, n_union]; float64 like numpy.

    recall = TP / sum(gt + 1e-6) (the 1e-6 is added per pixel before the sum, :86), precision = TP / (n_pred + 1e-6),
    f1 = 2pr / (p + r + 1e-6), iou = (TP + 1e-5) / (n_union + 1e-5)."""
    c = counts.to(torch.float64)
    tp, n_pred, n_gt, n_union = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    recall = tp / (n_gt + n_pixels * 1e-6)
    precision = tp / (n_pred + 1e-6)
    f1 = 2 * (precision * recall) / (precision + recall + 1e-6)
    iou = (tp + 1e-5) / (n_union + 1e-5)
    return f1, iou


def local_sums(counts: torch.Tensor, n_pixels: int) -> torch.Tensor:
    """[sum f1, sum iou, n_valid] over the clips measure.py keeps (f1 <= 1 and iou <= 1, :121)."""
    if counts.numel() == 0:
        return torch.zeros(3, dtype=torch.float64)
    f1, iou = f1_iou_from_counts(counts.cpu(), n_pixels)
    keep = (f1 <= 1) & (iou <= 1)
    return torch.stack([f1[keep].sum(), iou[keep].sum(), keep.sum().to(torch.float64)])


def reduce_means(sums: torch.Tensor, device=None):
    """All-reduce(SUM) of the 3 x fp64 partial sums -> (mean F1, mean IoU, n_valid).  Works without a process group."""
    t = sums.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    n = float(t[2])
    return (float(t[0]) / n if n else float("nan"), float(t[1]) / n if n else float("nan"), int(n))


def gather_count_table(counts: torch.Tensor, n_total: int, device=None):
    """Optional audit: all-gather the per-clip int64 count rows into the full (n_total, 4) table (rank order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts.cpu()
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width, 4), dtype=torch.int64, device=device or counts.device)
    pad[: counts.shape[0]] = counts.to(pad.device)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo].cpu() for o, (lo, hi) in zip(out, sizes)], 0)


class ShardedEvaluator:
    """Runs `predict(clips) -> logits (b,1,H,W)` over this rank's shard in micro-batches and reduces the metric.

    make_batch(lo, hi) -> (clips, gt) returns the clips [lo, hi) of the split (any device/dtype predict accepts) and
    their ground-truth masks (b,H,W) (bool/uint8), already on the prediction device."""

    def __init__(self, predict, make_batch, n_clips: int, micro_batch: int, counts_fn=None):
        self.predict = predict
        self.make_batch = make_batch
        self.n_clips = n_clips
        self.micro_batch = micro_batch
        self.counts_fn = counts_fn or self._device_counts

    @staticmethod
    def _device_counts(logits, gt):
        from . import ops
        return ops.mask_counts(logits.contiguous(), gt.to(torch.uint8).contiguous(), want_mask=False)[1]

    def run(self, device=None):
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        lo, hi = shard_bounds(self.n_clips, rank, world)
        rows, n_pixels = [], None
        for s in range(lo, hi, self.micro_batch):
            clips, gt = self.make_batch(s, min(hi, s + self.micro_batch))
            logits = self.predict(clips)
            n_pixels = logits.shape[-1] * logits.shape[-2]
            rows.append(self.counts_fn(logits, gt))
        counts = torch.cat(rows, 0) if rows else torch.zeros((0, 4), dtype=torch.int64)
        if counts.is_cuda:
            from . import ops
            counts = counts.cpu()                      # the shard's results are on the host: the range flag is final too
            ops.check_f16_range(device)                # 'fp16' mode left the half range -> raise, never report silently
        f1, iou, n = reduce_means(local_sums(counts, n_pixels or 1), device)
        return {"f1": f1, "iou": iou, "n_valid": n, "counts": counts, "shard": (lo, hi)}
