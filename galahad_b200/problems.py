"""Synthetic inputs of BASELINE.json's configs and the orderings we pass to
analyse.  The reference ships no ordering that works in a stock build except a user-supplied PERM
(METIS / MC68 are stubs), so every config passes a geometric nested-dissection
order explicitly; oracle and engine therefore always see identical input.

All matrices are returned the way SLS hands them to its back end (sls.f90:2853-2968): lower triangle
by columns (== upper by rows), row indices ascending, diagonal first and always present, 0-based.
"""
import numpy as np
import scipy.sparse as sp

SEED = 20261018


def lower_csc(A):
    """scipy symmetric matrix -> (n, ptr int64, row int32, val float64), diag present & first."""
    A = sp.csc_matrix(A)
    n = A.shape[0]
    Lo = sp.tril(A, format="csc")
    # make sure every diagonal entry is stored (explicit zero if absent), as SLS does
    d = Lo.diagonal()
    has = np.zeros(n, bool)
    coo = Lo.tocoo()
    has[coo.row[coo.row == coo.col]] = True
    if not has.all():
        miss = np.nonzero(~has)[0]
        Lo = (Lo + sp.csc_matrix((np.full(miss.size, 1.0), (miss, miss)), shape=(n, n))).tocsc()
        Lo.sort_indices()
        # subtract the placeholder back without dropping the stored entry
        for j in miss:
            Lo.data[Lo.indptr[j]] -= 1.0
    Lo.sort_indices()
    del d
    return n, Lo.indptr.astype(np.int64), Lo.indices.astype(np.int32), Lo.data.astype(np.float64)


def full_from_lower(n, ptr, row, val):
    col = np.repeat(np.arange(n), np.diff(ptr))
    Lo = sp.csc_matrix((val, (row, col)), shape=(n, n))
    return (Lo + sp.tril(Lo, -1).T).tocsr()


def grid_laplacian(nx, ny, nz, stencil=7, shift=0.0):
    """7-point (diag 6, off -1) or 27-point (diag 26, off -1) Dirichlet Laplacian + shift*I."""
    N = nx * ny * nz
    idx = np.arange(N).reshape(nz, ny, nx)
    rows, cols = [], []
    if stencil == 7:
        offs = [(0, 0, 1), (0, 1, 0), (1, 0, 0)]
        diag = 6.0
    else:
        offs = [(dz, dy, dx) for dz in (0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
                if (dz, dy, dx) > (0, 0, 0)]
        diag = 26.0
    for dz, dy, dx in offs:
        z0, z1 = max(0, -dz), nz - max(0, dz)
        y0, y1 = max(0, -dy), ny - max(0, dy)
        x0, x1 = max(0, -dx), nx - max(0, dx)
        a = idx[z0:z1, y0:y1, x0:x1].ravel()
        b = idx[z0 + dz:z1 + dz, y0 + dy:y1 + dy, x0 + dx:x1 + dx].ravel()
        rows.append(np.maximum(a, b))
        cols.append(np.minimum(a, b))
    rows = np.concatenate(rows + [np.arange(N)])
    cols = np.concatenate(cols + [np.arange(N)])
    vals = np.concatenate([np.full(rows.size - N, -1.0), np.full(N, diag + shift)])
    return sp.csc_matrix((vals, (rows, cols)), shape=(N, N))   # lower triangle only


def nd_order_sites(dims, leaf=8):
    """Geometric nested dissection of an nx*ny*nz grid of sites (x fastest).  Returns seq: the site
    indices in elimination order (children first, separator plane last)."""
    nx, ny, nz = dims
    idx = np.arange(nx * ny * nz).reshape(nz, ny, nx)
    out = []

    def rec(z0, z1, y0, y1, x0, x1):
        ext = (z1 - z0, y1 - y0, x1 - x0)
        if min(ext) <= 0:
            return
        if max(ext) <= leaf:
            out.append(idx[z0:z1, y0:y1, x0:x1].ravel())
            return
        ax = int(np.argmax(ext))
        lo = (z0, y0, x0)[ax]
        mid = lo + ext[ax] // 2
        if ax == 0:
            rec(z0, mid, y0, y1, x0, x1); rec(mid + 1, z1, y0, y1, x0, x1)
            out.append(idx[mid:mid + 1, y0:y1, x0:x1].ravel())
        elif ax == 1:
            rec(z0, z1, y0, mid, x0, x1); rec(z0, z1, mid + 1, y1, x0, x1)
            out.append(idx[z0:z1, mid:mid + 1, x0:x1].ravel())
        else:
            rec(z0, z1, y0, y1, x0, mid); rec(z0, z1, y0, y1, mid + 1, x1)
            out.append(idx[z0:z1, y0:y1, mid:mid + 1].ravel())

    rec(0, nz, 0, ny, 0, nx)
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def order_from_sequence(seq):
    order = np.empty(seq.size, np.int32)
    order[seq] = np.arange(seq.size, dtype=np.int32)
    return order


def laplace_problem(nx, ny=None, nz=None, stencil=7, shift=0.0, leaf=8):
    """Configs C2 / C5 (plain): SPD grid Laplacian with geometric ND order."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    A = grid_laplacian(nx, ny, nz, stencil, shift)
    n, ptr, row, val = lower_csc(A)
    order = order_from_sequence(nd_order_sites((nx, ny, nz), leaf))
    return dict(n=n, ptr=ptr, row=row, val=val, order=order, inertia=(n, 0, 0), name=f"lap{stencil}_{nx}x{ny}x{nz}")


def kkt_grid_problem(nx, ny=None, nz=None, stencil=7, seed=SEED, dlo=1e-6, dhi=1.0, leaf=8):
    """Config C3 / C5 (KKT): K = [H A^T; A -D], H = grid Laplacian + 0.1 I on nx*ny*nz (nz even),
    A row i couples x_i and x_{i+m} (m = n/2) with +1,+1, D log-uniform in [dlo,dhi].
    Inertia is exactly (n, m, 0).  Ordering: the grid is folded in z so that x_i, x_{i+m} and their
    constraint share a site; sites are ordered by geometric ND, the constraint after its two ends."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    assert nz % 2 == 0
    n = nx * ny * nz
    m = n // 2
    H = grid_laplacian(nx, ny, nz, stencil, 0.1)
    rng = np.random.default_rng(seed)
    D = np.exp(rng.uniform(np.log(dlo), np.log(dhi), m))
    i = np.arange(m)
    Arows = np.concatenate([n + i, n + i])
    Acols = np.concatenate([i, i + m])
    Avals = np.ones(2 * m)
    Dm = sp.csc_matrix((-D, (n + i, n + i)), shape=(n + m, n + m))
    Hbig = sp.csc_matrix((H.data, H.indices, np.concatenate([H.indptr, np.full(m, H.indptr[-1])])), shape=(n + m, n + m))
    K = Hbig + sp.csc_matrix((Avals, (Arows, Acols)), shape=(n + m, n + m)) + Dm
    N, ptr, row, val = lower_csc(K)
    seq_sites = nd_order_sites((nx, ny, nz // 2), leaf)
    seq = np.empty(3 * m, np.int64)
    seq[0::3] = seq_sites
    seq[1::3] = seq_sites + m
    seq[2::3] = n + seq_sites
    order = order_from_sequence(seq)
    return dict(n=N, ptr=ptr, row=row, val=val, order=order, inertia=(n, m, 0), name=f"kkt{stencil}_{nx}x{ny}x{nz}")


def qpband_problem(N, seed=SEED, it=0, leaf=64):
    """Config C4: K = [H + S  A^T; A  -C] for QPBAND (examples/QPBAND.SIF:5-67): H = tridiag(-1,2,-1),
    m = N/2 constraints x_i + x_{m+i}; S, C synthetic barrier diagonals, log-uniform 1e-8..1e8 (S)
    and 1e-8..1 (C), reseeded per interior-point iteration `it`.  Inertia (N, m, 0)."""
    m = N // 2
    rng = np.random.default_rng(seed + 7919 * it)
    S = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), N))
    Cd = np.exp(rng.uniform(np.log(1e-8), np.log(1.0), m))
    i = np.arange(N)
    rows = np.concatenate([i, i[1:], N + np.arange(m), N + np.arange(m), N + np.arange(m)])
    cols = np.concatenate([i, i[:-1], np.arange(m), m + np.arange(m), N + np.arange(m)])
    vals = np.concatenate([2.0 + S, np.full(N - 1, -1.0), np.ones(m), np.ones(m), -Cd])
    K = sp.csc_matrix((vals, (rows, cols)), shape=(N + m, N + m))
    n, ptr, row, val = lower_csc(K)
    # fold the chain: site s holds x_s, x_{m+s} and constraint s; 1-D ND over sites, the two chain
    # ends (coupled through x_{m-1} -- x_m) go last
    seq_sites = nd_order_sites((m, 1, 1), leaf)
    ends = np.array([0, m - 1]) if m > 1 else np.array([0])
    seq_sites = np.concatenate([seq_sites[~np.isin(seq_sites, ends)], ends])
    seq = np.empty(3 * m, np.int64)
    seq[0::3] = seq_sites
    seq[1::3] = seq_sites + m
    seq[2::3] = N + seq_sites
    order = order_from_sequence(seq)
    return dict(n=n, ptr=ptr, row=row, val=val, order=order, inertia=(N, m, 0), name=f"qpband_{N}_{it}")


def random_sym_problem(n, density=0.05, kind="indef", seed=SEED):
    """Small random test matrices: 'spd', 'indef' (symmetric, mixed signs) or 'kkt'."""
    rng = np.random.default_rng(seed)
    if kind == "kkt":
        nh = (2 * n) // 3
        mc = n - nh
        B = sp.random(nh, nh, density=density, random_state=rng, data_rvs=rng.standard_normal)
        H = (B @ B.T + sp.identity(nh) * 0.5).tocsc()
        A = sp.random(mc, nh, density=min(1.0, 3 * density), random_state=rng, data_rvs=rng.standard_normal).tocsc()
        A = A + sp.csc_matrix((np.ones(mc), (np.arange(mc), rng.permutation(nh)[:mc])), shape=(mc, nh))
        Dd = sp.diags(-np.exp(rng.uniform(np.log(1e-8), 0, mc)))
        K = sp.bmat([[H, A.T], [A, Dd]]).tocsc()
        inertia = None
    else:
        B = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal)
        S = (B + B.T).tocsc()
        if kind == "spd":
            K = (S @ S.T + sp.identity(n)).tocsc()
            inertia = (n, 0, 0)
        else:
            K = (S + sp.diags(rng.standard_normal(n) * 0.1)).tocsc()
            inertia = None
    N, ptr, row, val = lower_csc(K)
    order = rng.permutation(N).astype(np.int32)
    return dict(n=N, ptr=ptr, row=row, val=val, order=order, inertia=inertia, name=f"rand_{kind}_{n}")


def backward_error(prob, x, b):
    """Scaled backward error ||Ax-b||_inf / (||A||_inf ||x||_inf + ||b||_inf)."""
    A = full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"])
    r = A @ x - b
    an = abs(A).sum(axis=1).max()
    return float(np.abs(r).max() / (an * np.abs(x).max() + np.abs(b).max()))


def dense_front_problem(n, kind, seed=SEED):
    """A dense symmetric matrix factorised as ONE front (identity order).  kinds: 'spd', 'indef' (random),
    'zdiag' (zero diagonal: 2x2 pivots everywhere), 'kkt' (zero (2,2) block), 'small' (badly scaled).
    Returns (A, problem dict); the expected inertia comes from the eigenvalues."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (n, n))
    A = (A + A.T) / 2
    if kind == "zdiag":
        np.fill_diagonal(A, 0.0)
    elif kind == "kkt":
        k = n // 3
        A[n - k:, n - k:] = 0.0
        A[:n - k, :n - k] += np.eye(n - k) * 2
    elif kind == "spd":
        A = A @ A.T + n * np.eye(n)
    elif kind == "small":
        A[:, : n // 2] *= 1e-3
        A[: n // 2, :] *= 1e-3
    il = np.tril_indices(n)
    order = np.lexsort((il[0], il[1]))           # column-major lower triangle
    row, col = il[0][order], il[1][order]
    ptr = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=n))]).astype(np.int64)
    val = A[row, col].copy()
    neg = int((np.linalg.eigvalsh(A) < 0).sum())
    return A, dict(n=n, ptr=ptr, row=row.astype(np.int32), val=val, order=np.arange(n, dtype=np.int32),
                   inertia=(n - neg, neg, 0), name=f"dense_{kind}_{n}")
