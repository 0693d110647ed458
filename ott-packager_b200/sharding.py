"""Channel -> GPU placement for a multi-GPU box: channels are independent (each frame depends only on
its own channel's previous/next frame), so the path shards by channel with NO data-path collective.  A channel never
spans GPUs.  Heavier channels (4K) are placed first so every GPU gets its share of them (BASELINE config 5)."""


def channel_cost(spec):
    """relative per-frame cost of a channel: source pixels x frame rate (4K60 = 8x a 1080i29.97 channel)"""
    return spec["w"] * spec["h"] * spec.get("fps", 30.0)


def shard_channels(specs, world_size):
    """specs: list of dicts with at least w, h (optionally fps).  Returns a list (one per rank) of channel indices.
    Greedy longest-processing-time placement: deterministic, every rank computes the same answer without talking."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(specs)), key=lambda i: (-channel_cost(specs[i]), i))
    load = [0.0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += channel_cost(specs[i])
    for r in range(world_size):
        out[r].sort()
    return out


def my_channels(specs, rank, world_size):
    return shard_channels(specs, world_size)[rank]
