'''
Host-side helpers for the two multi-GPU modes of the hot path (one process per
GPU, ``torch.distributed``):

  * training is data parallel: every rank holds a replica and a slice of the
    batch; the flat fp32 gradient arena is summed over ranks once per step, in
    buckets that follow the order in which backward finishes them;
  * generation shards independent sequences over ranks with no communication.

The reference has no distributed code at all; these are
additions around its single-process loop (composer/models/transformer.py:907-946).
Pure Python on purpose: testable on CPU with the ``gloo`` backend.
'''


def gradient_buckets(layout, decoder_layers_count):
    '''
    ``layout`` maps Keras variable names to ``(offset, rows, cols)`` in the flat
    arena.  Returns ``[(start, end), ...]`` element ranges in backward-completion
    order: ln_f, decoder blocks L..1, then the embeddings (wte also receives the
    tied-logits gradient in the first stage, so it must go last).  The ranges are
    disjoint and cover every variable.
    '''

    def span(first, last):
        start = layout[first][0]
        offset, rows, cols = layout[last]
        return start, offset + rows * cols

    buckets = [span('ln_f/gamma', 'ln_f/beta')]
    for layer in range(decoder_layers_count, 0, -1):
        buckets.append(span('h_%d/ln_1/gamma' % layer, 'h_%d/mlp/c_proj/bias' % layer))
    buckets.append(span('wte/weight', 'wpe/embeddings'))
    return buckets


def allreduce_buckets(flat, buckets, group=None, after_bucket=None):
    '''
    Sum-all-reduces ``flat[start:end]`` for every bucket asynchronously, calling
    ``after_bucket(index)`` *before* each launch (the hook enqueues the backward
    stage that produces the bucket), then waits for all of them.
    '''

    import torch.distributed as dist

    pending = []
    for index, (start, end) in enumerate(buckets):
        if after_bucket is not None:
            after_bucket(index)
        pending.append(dist.all_reduce(flat[start:end], group=group, async_op=True))
    for work in pending:
        work.wait()


def shard_range(total, world_size, rank):
    '''Contiguous block of ``range(total)`` owned by ``rank`` (blocks differ by at most one).'''

    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def agree_on_seed(seed, group=None):
    '''
    Every rank adopts rank 0's seed.  Without ``--seed`` each process derives its own seed from the clock
    (composer/cli.py:51-57), and that seed drives the initial weights, the file shuffle and the window order:
    replicas that disagree on it never train one model.  No-op for a single process.
    '''

    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(seed)
    holder = [int(seed)]
    dist.broadcast_object_list(holder, src=0, group=group)
    return int(holder[0])


def rank_dropout_seed(seed, rank):
    '''The dropout key of a data-parallel rank: masks must differ between ranks, everything else must not.'''

    return (int(seed) + 0x9E3779B97F4A7C15 * int(rank)) % (1 << 64)
