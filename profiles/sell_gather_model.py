"""CPU model of the gather traffic of the Kronecker SELL kernel (no GPU needed).

For the scalar operator S of A00 = S (x) I_3 on the bench lattice, builds the SELL-32-sigma layout the
library builds (rows sorted by length inside windows of sigma rows, stable) and counts, per gather
request of a warp (the k-th entries of the 32 rows of a slice), the distinct 32-byte sectors and
128-byte lines touched by the 24-byte groups x[3c .. 3c+2] -- ncu's
l1tex__t_sectors / l1tex__t_requests and the LSU wavefront count are proportional to these.
Compares the caller's lattice numbering with alternatives a library-internal renumbering could use.
    python profiles/sell_gather_model.py [n1]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lattice_operator(n):
    """Pattern of the scalar P2 operator on n^3 bricks x 6 Kuhn tetrahedra, lattice numbering
    (x fastest) of the (2n+1)^3 nodes -- the numbering bench_inputs.py uses."""
    m = 2 * n + 1
    idx = np.arange(m ** 3).reshape(m, m, m)           # [z, y, x]
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    rows, cols = [], []
    bz, by, bx = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = np.stack([bx.ravel(), by.ravel(), bz.ravel()], 1) * 2        # lattice coords of brick corner
    for p in perms:
        v = [np.zeros(3, dtype=int)]
        for ax in p:
            e = v[-1].copy()
            e[ax] = 2
            v.append(e)
        v = np.array(v)                                                  # 4 vertices (lattice offsets)
        nodes = [v[i] for i in range(4)] + [(v[i] + v[j]) // 2 for i in range(4) for j in range(i + 1, 4)]
        nodes = np.array(nodes)                                          # 10 P2 nodes of the tet
        g = base[:, None, :] + nodes[None, :, :]                         # [cells, 10, 3]
        gid = idx[g[..., 2], g[..., 1], g[..., 0]]
        rows.append(np.repeat(gid, 10, axis=1).ravel())
        cols.append(np.tile(gid, (1, 10)).ravel())
    r, c = np.concatenate(rows), np.concatenate(cols)
    S = sp.csr_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(m ** 3, m ** 3))
    S.sum_duplicates()
    S.sort_indices()
    return S, m


def sell_requests(S, sigma):
    """Yield, per slice, the [len, 32] array of column ids (padding = first column of the row)."""
    n = S.shape[0]
    lens = np.diff(S.indptr)
    perm = np.arange(n)
    for w0 in range(0, n, sigma):
        w = perm[w0:w0 + sigma]
        perm[w0:w0 + sigma] = w[np.argsort(-lens[w], kind="stable")]
    pad = (-n) % 32
    perm = np.concatenate([perm, np.full(pad, -1)])
    total_pad = 0
    for s0 in range(0, perm.size, 32):
        rows = perm[s0:s0 + 32]
        L = max(lens[r] for r in rows if r >= 0)
        block = np.empty((L, 32), dtype=np.int64)
        for l, r in enumerate(rows):
            if r < 0:
                block[:, l] = 0
                continue
            cs = S.indices[S.indptr[r]:S.indptr[r + 1]]
            block[:cs.size, l] = cs
            block[cs.size:, l] = cs[0]
            total_pad += L - cs.size
        yield block, total_pad


def count(S, sigma, bytes_per_node=24, stride_nodes=1, sample=4000):
    """Mean distinct sectors / lines per gather request and padding fraction.  bytes_per_node=24:
    interleaved components (AoS); 8: one component array (SoA, counted once per component)."""
    sect, line, nreq, pad = 0, 0, 0, 0
    rng = np.random.default_rng(0)
    nsl = (S.shape[0] + 31) // 32
    pick = set(rng.choice(nsl, size=min(sample, nsl), replace=False).tolist())
    for i, (block, total_pad) in enumerate(sell_requests(S, sigma)):
        pad = total_pad
        if i not in pick:
            continue
        a0 = block * bytes_per_node
        a1 = a0 + bytes_per_node - 1
        for k in range(block.shape[0]):
            s = np.union1d(a0[k] // 32, a1[k] // 32)
            if bytes_per_node == 24:
                s = np.union1d(s, (a0[k] + 12) // 32)
            sect += s.size
            line += np.unique(s // 4).size
            nreq += 1
    return sect / nreq, line / nreq, pad / (S.nnz + pad)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S, m = lattice_operator(n)
    lens = np.diff(S.indptr)
    print(f"n = {n}: {S.shape[0]} nodes, {S.nnz} stored entries, row lengths {sorted(set(lens.tolist()))}")
    z, y, x = np.unravel_index(np.arange(m ** 3), (m, m, m))
    cls = (x % 2) + 2 * (y % 2) + 4 * (z % 2)
    by_class = np.argsort(cls, kind="stable")                       # parity-class-major numbering
    inv = np.empty_like(by_class)
    inv[by_class] = np.arange(by_class.size)
    Sc = S[by_class][:, by_class].tocsr()
    Sc.sort_indices()
    by_len = np.argsort(-lens, kind="stable")                       # global sort by row length
    Sl = S[by_len][:, by_len].tocsr()
    Sl.sort_indices()
    def window_sorted(M, sigma):
        """Symmetric permutation by the SELL row order itself: rows sorted by length inside
        windows of sigma nodes become the new numbering (mesh-agnostic)."""
        n_ = M.shape[0]
        ln = np.diff(M.indptr)
        perm = np.arange(n_)
        for w0 in range(0, n_, sigma):
            w = perm[w0:w0 + sigma]
            perm[w0:w0 + sigma] = w[np.argsort(-ln[w], kind="stable")]
        P = M[perm][:, perm].tocsr()
        P.sort_indices()
        return P
    print("layout                                   | sectors/request | lines/request | padding")
    for name, M, sigma, bpn in (("lattice numbering, sigma 1024 (library)", S, 1024, 24),
                                ("lattice numbering, sigma 32 (no sorting)", S, 32, 24),
                                ("lattice numbering, sigma 8192", S, 8192, 24),
                                ("parity-class-major numbering, sigma 1024", Sc, 1024, 24),
                                ("parity-class-major numbering, sigma 32", Sc, 32, 24),
                                ("row-length-major numbering, sigma 1024", Sl, 1024, 24),
                                ("window-sorted numbering (1024), sigma 32", window_sorted(S, 1024), 32, 24),
                                ("window-sorted numbering (8192), sigma 32", window_sorted(S, 8192), 32, 24),
                                ("window-sorted numbering (65536), sigma 32", window_sorted(S, 65536), 32, 24),
                                ("lattice numbering, component arrays (x3)", S, 1024, 8),
                                ("parity-class-major, component arrays (x3)", Sc, 1024, 8)):
        s, l, p = count(M, sigma, bpn)
        mult = 3 if bpn == 8 else 1
        print(f"{name:41s}| {s * mult:15.1f} | {l * mult:13.1f} | {100 * p:5.1f} %")


if __name__ == "__main__":
    main()
