"""Shard arithmetic shared by compress_multi callers and bench.py (one process per GPU).

``get_range`` is the reference's split rule (src/enc/threading/mod.rs:333-335).  A shard sees the previous 2^lgwin
bytes of input as its window and ends byte aligned unless it is the last one, so the shard outputs concatenate with
a plain copy (the reference needs BroCatli's bit-shifting stitcher for the same job, src/concat/mod.rs:331-449)."""


def get_range(index: int, num_shards: int, total: int):
    return index * total // num_shards, (index + 1) * total // num_shards


def shard_flags(index: int, num_shards: int):
    """(first, last, byte_align) flags of b200_encoder_compress_range for shard `index`."""
    return index == 0, index + 1 == num_shards, index + 1 != num_shards


def concat_shards(parts):
    return b"".join(parts)
