"""Row sharding of one minibatch over the ranks of a box: contiguous row ranges, all S samples of a
row stay on the rank that owns the row; parameters (and the Kuu factors) are replicated; one all-reduce(sum) of the
flat [elbo | gradients] buffer per evaluation, with the KL term weighted 1/world_size on every rank."""


def shard_range(num_rows, rank, world_size):
    """Rows [lo, hi) of a minibatch owned by `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (num_rows * rank) // world_size, (num_rows * (rank + 1)) // world_size


def kl_weight(world_size):
    """Share of the (replicated) KL term a rank adds before the all-reduce; 1/2, 1/4, 1/8 are exact in binary."""
    return 1.0 / float(world_size)


def philox_row_id(row_offset, sample, local_row, sample_stride):
    """Global row id that keys the Philox draw of (sample, local row): identical for any sharding of the rows."""
    return row_offset + sample * sample_stride + local_row


def nccl_options(max_ctas=4):
    """ProcessGroupNCCL options for the process group handed to DSDGP.distribute(): at most `max_ctas` CTAs per collective.
    The per-layer gradient buckets (a few MB) are all-reduced on a communication stream WHILE the persistent row kernels of
    the earlier layers hold every SM with one 220 KB CTA each; NCCL's default CTA count made every other T step 0.4 ms slower
    on 2, 4 and 8 GPUs (9.8 / 10.2 ms bimodal), because the row-kernel CTAs of the SMs NCCL reached first had to wait for it.
    Four CTAs move these buckets over NVLink just as fast and leave the step at its single-GPU time."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    try:
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = 1
    except AttributeError:        # a torch without ncclConfig_t in its bindings: NCCL's defaults
        pass
    return opts


class BucketReducer:
    """Sum-all-reduce of the flat [elbo | gradients] buffer in per-layer buckets, each issued as soon as that layer's
    gradients are final (last layer first), so that the collectives run on NCCL's stream beside the backward kernels of
    the earlier layers instead of after the whole evaluation.  `bucket()` is called with the stream that produced the slice
    current (torch.distributed makes NCCL's stream wait for it); `finish()` makes the current stream wait for every
    bucket.  Calling the object reduces a whole buffer in one collective (ELBO-only evaluations)."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group
        self.works = []

    def bucket(self, tensor):
        self.works.append(self.dist.all_reduce(tensor, group=self.group, async_op=True))

    def finish(self):
        for w in self.works:
            w.wait()
        self.works.clear()

    def __call__(self, tensor):
        self.dist.all_reduce(tensor, group=self.group)
