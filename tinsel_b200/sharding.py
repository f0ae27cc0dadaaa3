"""Host-side image-plane sharding rule shared by bench.py and the tests.

Rank r of n owns the 4-pixel-high tile rows t with t % n == r (tb200_set_shard; the device side is
decode_sample() in csrc/kernels.cu).  Every sample belongs to exactly one shard, so the sum of the
per-rank accumulators is the unsharded image (up to fp32 summation order)."""

TILE_ROWS = 4


def shard_rows(height, shard, num_shards):
    """Pixel rows whose samples shard `shard` of `num_shards` traces."""
    rows = []
    t = shard
    while t * TILE_ROWS < height:
        rows.extend(range(t * TILE_ROWS, min(height, (t + 1) * TILE_ROWS)))
        t += num_shards
    return rows


def shard_sample_count(width, height, shard, num_shards, spp=1):
    return len(shard_rows(height, shard, num_shards)) * width * spp
