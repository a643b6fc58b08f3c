"""Static round-robin of chunk tiles over devices: chunks are independent, so the
multi-GPU path has no data-path collective; tile t belongs to rank t mod world_size and results are
gathered by the host.  Pure host logic (tested with gloo, world_size 2, on CPU)."""


def tiles(n_chunks, tile_chunks):
    """[(first_chunk, n)] covering [0, n_chunks) in tiles of at most tile_chunks."""
    return [(c0, min(tile_chunks, n_chunks - c0)) for c0 in range(0, n_chunks, tile_chunks)]


def tile_owner(tile_index, world_size):
    return tile_index % world_size


def tiles_for_rank(n_chunks, tile_chunks, rank, world_size):
    return [(i, c0, n) for i, (c0, n) in enumerate(tiles(n_chunks, tile_chunks)) if tile_owner(i, world_size) == rank]


def gather_order(n_chunks, tile_chunks, world_size):
    """For every tile, (owner rank, index within that rank's list): how the host reassembles results."""
    seen = [0] * world_size
    out = []
    for i, _ in enumerate(tiles(n_chunks, tile_chunks)):
        r = tile_owner(i, world_size)
        out.append((r, seen[r]))
        seen[r] += 1
    return out
