"""Frame-level data parallelism: frames are independent units, so a batch is
dealt to ranks without any data-path collective.  The only shared state is the LUT blob, broadcast
once from rank 0."""


def frames_for_rank(n_frames, rank, world):
    """contiguous chunks (config 4: 256 frames -> 32 per GPU); remainder spread over the first ranks"""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def broadcast_lut_blob(blob_tensor, dist, src=0):
    """one collective for the whole job: rank `src` built the blob with the reference's host
    expressions; every other rank receives a bit-identical copy"""
    dist.broadcast(blob_tensor, src=src)
    return blob_tensor
