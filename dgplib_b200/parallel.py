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


def nccl_options(world_size=None, max_ctas=None):
    """ProcessGroupNCCL options for the process group handed to DSDGP.distribute().

    The per-layer gradient buckets (a few MB) are all-reduced on a communication stream WHILE the persistent row kernels of
    the earlier layers hold every SM with one 220 KB CTA each, so the number of CTAs NCCL may use is a trade: few CTAs leave
    the row kernels alone, many finish the (partly exposed) last bucket sooner.  Measured on T (10 k rows per GPU):
    2 ranks, NCCL default: steps bimodal 9.8 / 10.2 ms, mean 10.02; max_ctas = 4: every step 9.85-9.97 ms, mean 9.90.
    4 ranks, default: 10.16 ms; max_ctas = 4: 10.32 ms (the ring over four GPUs needs more than four CTAs to hide).
    Hence: four CTAs for two ranks, NCCL's own choice otherwise, unless `max_ctas` says so."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    if max_ctas is None and world_size == 2:
        max_ctas = 4
    if max_ctas is not None:
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
