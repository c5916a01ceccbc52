"""Process-grid arithmetic of the multi-GPU paths (pure host logic, no device code).

GEMM (BASELINE configs[1]): 2D output-tile sharding -- rank (r, c) of a pr x pc grid owns
C[r-block, c-block] and needs the row panel A[r-block, :] and the column panel B[:, c-block]
(SURVEY.md §8e).  Cholesky / LU at large N: 1D block-cyclic column blocks.
"""
from __future__ import annotations

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


def process_grid(world: int) -> tuple[int, int]:
    if world in GRIDS:
        return GRIDS[world]
    pr = 1
    while pr * pr * 2 <= world and world % (pr * 2) == 0:
        pr *= 2
    return pr, world // pr


def split(n: int, parts: int, i: int) -> tuple[int, int]:
    """[begin, end) of part i when n is cut into `parts` nearly equal contiguous pieces."""
    base, rem = divmod(n, parts)
    b = i * base + min(i, rem)
    return b, b + base + (1 if i < rem else 0)


def gemm_tile(rank: int, world: int, m: int, n: int):
    """(row_begin, row_end, col_begin, col_end) of the C tile owned by `rank`."""
    pr, pc = process_grid(world)
    r, c = divmod(rank, pc)
    r0, r1 = split(m, pr, r)
    c0, c1 = split(n, pc, c)
    return r0, r1, c0, c1


def block_cyclic_owner(block: int, world: int) -> int:
    """Owner of column block `block` in the 1D block-cyclic layout of the factorizations."""
    return block % world


def block_cyclic_local_blocks(rank: int, world: int, nblocks: int) -> list[int]:
    return list(range(rank, nblocks, world))


def gemm_piece_class(t: int, kp: int, kca: int, kcb: int, my_r: int, my_c: int) -> int:
    """Which operands of K-piece t (k in [t*kp, (t+1)*kp)) rank (my_r, my_c) already owns in the non-replicated
    layout (A row panel cut into K-chunks of kca columns, chunk c owned by grid column c; B column panel cut into
    K-chunks of kcb rows, chunk r owned by grid row r): 0 both, 1 only A (B arrives with the column gather),
    2 only B (A arrives with the row gather), 3 neither."""
    a_loc, b_loc = (t * kp) // kca == my_c, (t * kp) // kcb == my_r
    return 0 if (a_loc and b_loc) else (1 if a_loc else (2 if b_loc else 3))


def gemm_piece_order(n_k: int, kp: int, kca: int, kcb: int, my_r: int, my_c: int) -> list[int]:
    """Order in which a rank accumulates the K-pieces of its C tile: local pieces first, then those that wait only
    for the (smaller, earlier) B gather, then those that wait for the A gather, then both -- so the DMMA work
    overlaps the exchange."""
    assert kca % kp == 0 and kcb % kp == 0 and n_k % kp == 0
    return sorted(range(n_k // kp), key=lambda t: (gemm_piece_class(t, kp, kca, kcb, my_r, my_c), t))

