"""Symbolic-analysis oracle: an independent restatement, in plain Python, of the reference's analyse path

    check_order / expand_pattern / build_map          src/ssids/anal.f90:147-197, 37-80, 1129-1231
    basic_analyse and its children                    src/spral/core_analyse.f90:38-1098
    analyse_phase (sequence, invp, order of null cols) src/ssids/anal.f90:940-1013

TEST INFRASTRUCTURE ONLY (tests/ import it; the product path is galahad_b200/csrc/analyse.cpp).  It follows the Fortran
statement by statement with 1-based arrays (index 0 unused) so that tie-breaking -- the order children come off their
linked lists, the stable insertion / merge sorts of find_supernodes, the depth-first numbering of find_postorder, the
double-transpose sort of the row lists -- is the reference's.  The Fortran itself cannot be compiled in this image (no
Fortran compiler): this restatement is the second, independently written opinion that the C++ one is compared with,
entry by entry (`parity unpinned` against SSIDS's own output in the strict sense).

Pure-Python loops: meant for n up to a few 10^4.
"""
import numpy as np

MINSZ_MS = 16            # core_analyse.f90:20
HUGE = (1 << 63) - 1     # huge(ezero), integer(long)


def check_order(n, order):
    """anal.f90:147-197.  order: 1-based positions (list, index 0 unused).  Returns invp or None (SSIDS_ERROR_ORDER)."""
    invp = [0] * (n + 1)
    for i in range(1, n + 1):
        order[i] = abs(order[i])
    i = 1
    while i <= n:
        j = order[i]
        if j <= 0 or j > n:
            break
        if invp[j] != 0:
            break
        invp[j] = i
        i += 1
    if i - 1 != n:
        return None
    return invp


def expand_pattern(n, ptr, row):
    """anal.f90:37-80: lower CSC -> full pattern.  1-based ptr (n+2 entries, index 0 unused), row (index 0 unused)."""
    nz = ptr[n + 1] - 1
    aptr = [0] * (n + 2)
    arow = [0] * (2 * nz + 1)
    for j in range(1, n + 1):
        for kk in range(ptr[j], ptr[j + 1]):
            i = row[kk]
            aptr[i] += 1
            if j == i:
                continue
            aptr[j] += 1
    for j in range(2, n + 1):
        aptr[j] = aptr[j - 1] + aptr[j]
    aptr[n + 1] = aptr[n] + 1
    for j in range(1, n + 1):
        for kk in range(ptr[j], ptr[j + 1]):
            i = row[kk]
            arow[aptr[i]] = j
            aptr[i] -= 1
            if j == i:
                continue
            arow[aptr[j]] = i
            aptr[j] -= 1
    for j in range(1, n + 1):
        aptr[j] += 1
    return aptr, arow


def find_etree(n, ptr, row, perm, invp):
    """core_analyse.f90:173-223 (Liu's algorithm with path compression through a virtual forest)."""
    parent = [0] * (n + 1)
    vforest = [n + 1] * (n + 1)
    for piv in range(1, n + 1):
        rowidx = invp[piv]
        for i in range(ptr[rowidx], ptr[rowidx + 1]):
            j = perm[row[i]]
            if j >= piv:
                continue
            k = j
            while vforest[k] < piv:
                l = vforest[k]
                vforest[k] = piv
                k = l
            if vforest[k] == piv:
                continue
            parent[k] = piv
            vforest[k] = piv
        parent[piv] = n + 1
    return parent


def find_postorder(n, ptr, perm, invp, parent):
    """core_analyse.f90:233-352.  Relabels perm, invp, parent in place; returns realn."""
    realn = n
    chead = [-1] * (n + 2)
    cnext = [0] * (n + 2)
    for i in range(n, 0, -1):
        j = parent[i]
        cnext[i] = chead[j]
        chead[j] = i
    map_ = [0] * (n + 2)
    stack = [0] * (n + 2)
    shead = 1
    stack[shead] = n + 1
    id_ = n + 1
    while shead != 0:
        node = stack[shead]
        shead -= 1
        map_[node] = id_
        id_ -= 1
        if node == n + 1:
            i = chead[node]
            while i != -1:
                if ptr[invp[i] + 1] - ptr[invp[i]] == 0:
                    i = cnext[i]
                    continue
                shead += 1
                stack[shead] = i
                i = cnext[i]
            i = chead[node]
            while i != -1:
                if ptr[invp[i] + 1] - ptr[invp[i]] != 0:
                    i = cnext[i]
                    continue
                realn -= 1
                shead += 1
                stack[shead] = i
                i = cnext[i]
        else:
            i = chead[node]
            while i != -1:
                shead += 1
                stack[shead] = i
                i = cnext[i]
    tmp = invp[:]
    for i in range(1, n + 1):
        invp[map_[i]] = tmp[i]
    for i in range(1, n + 1):
        perm[invp[i]] = i
    tmp = [0] * (n + 1)
    for i in range(1, n + 1):
        tmp[i] = map_[parent[i]]
    for i in range(1, n + 1):
        parent[map_[i]] = tmp[i]
    return realn


def _find(vforest, u):
    """core_analyse.f90:505-521."""
    prev = -1
    current = u
    while vforest[current] != 0:
        prev = current
        current = vforest[current]
        if vforest[current] != 0:
            vforest[prev] = vforest[current]
    return current


def find_col_counts(n, ptr, row, perm, invp, parent):
    """core_analyse.f90:387-501 (Gilbert, Ng, Peyton)."""
    cc = [0] * (n + 2)
    first = list(range(n + 2))
    for i in range(1, n + 1):
        par = parent[i]
        first[par] = min(first[i], first[par])
        cc[i] = 1 if first[i] == i else 0
    cc[n + 1] = n + 1
    vforest = [0] * (n + 2)
    last_p = [0] * (n + 2)
    last_nbr = [0] * (n + 2)
    for piv in range(1, n + 1):
        col = invp[piv]
        for ii in range(ptr[col], ptr[col + 1]):
            u = perm[row[ii]]
            if u <= piv:
                continue
            if first[piv] > last_nbr[u]:
                cc[piv] += 1
                pp = last_p[u]
                if pp != 0:
                    lca = _find(vforest, pp)
                    cc[lca] -= 1
                last_p[u] = piv
            last_nbr[u] = piv
        par = parent[piv]
        cc[par] = cc[par] + cc[piv] - 1
        vforest[piv] = par
    return cc


def _sort_by_val(idx, lo, n, val):
    """core_analyse.f90:712-743: idx[lo+1 .. lo+n] into decreasing val (stable insertion from the back)."""
    if n >= MINSZ_MS:
        _sort_by_val_ms(idx, lo, n, val)
        return
    kor = n
    for _ in range(2, n + 1):
        ice_idx = idx[lo + kor - 1]
        ice_val = val[ice_idx]
        k = kor
        while k <= n:
            ik_idx = idx[lo + k]
            ik_val = val[ik_idx]
            if ice_val >= ik_val:
                break
            idx[lo + k - 1] = ik_idx
            k += 1
        idx[lo + k - 1] = ice_idx
        kor -= 1


def _sort_by_val_ms(idx, lo, n, val):
    """core_analyse.f90:750-801."""
    if n <= 1:
        return
    if n < MINSZ_MS:
        _sort_by_val(idx, lo, n, val)
        return
    mid = (n - 1) // 2 + 1
    _sort_by_val_ms(idx, lo, mid, val)
    _sort_by_val_ms(idx, lo + mid, n - mid, val)
    work = [0] + idx[lo + 1: lo + mid + 1]
    j = 1
    k = mid + 1
    jj = work[j]
    jj2 = val[jj]
    kk = idx[lo + k]
    kk2 = val[kk]
    i = 1
    while i <= n:
        if jj2 >= kk2:
            idx[lo + i] = jj
            j += 1
            if j > mid:
                break
            jj = work[j]
            jj2 = val[jj]
        else:
            idx[lo + i] = kk
            k += 1
            if k > n:
                break
            kk = idx[lo + k]
            kk2 = val[kk]
        i += 1
    if j <= mid:
        # Fortran: idx(i+1:n) = work(j:mid) with i the value at the exit
        idx[lo + i + 1: lo + n + 1] = work[j: mid + 1]


def find_supernodes(n, realn, parent, cc, nemin):
    """core_analyse.f90:536-707 (do_merge :806-822, merge_nodes :827-853)."""
    nelim = [1] * (n + 2)
    nvert = [1] * (n + 2)
    vhead = [-1] * (n + 2)
    vnext = [-1] * (n + 2)
    height = [1] * (n + 2)
    mark = [False] * (n + 2)
    ezero = [0] * (n + 2)
    ezero[n + 1] = HUGE
    totalwt = n
    nelim[n + 1] = totalwt + 1 + nemin
    chead = [-1] * (n + 2)
    cnext = [0] * (n + 2)
    for i in range(realn, 0, -1):
        j = parent[i]
        cnext[i] = chead[j]
        chead[j] = i
    child = [0] * (n + 2)
    for par in range(1, n + 2):
        nchild = 0
        node = chead[par]
        while node != -1:
            nchild += 1
            child[nchild] = node
            node = cnext[node]
        _sort_by_val(child, 0, nchild, cc)
        for j in range(1, nchild + 1):
            node = child[j]
            if ezero[par] == HUGE:
                merge = False
            else:
                merge = (cc[par] == cc[node] - 1 and nelim[par] == 1) or (nelim[par] < nemin and nelim[node] < nemin)
            if merge:
                vnext[node] = vhead[par]
                vhead[par] = node
                ezero[par] = ezero[par] + ezero[node] + (cc[par] - 1 + nelim[par] - cc[node] + 1) * nelim[par]
                nelim[par] += nelim[node]
                nvert[par] += nvert[node]
                height[par] = max(height[par], height[node])
                mark[node] = False
            else:
                mark[node] = True
    sperm = [0] * (n + 1)
    map_ = [0] * (n + 2)
    npar = [0] * (n + 2)
    sptr = [0] * (n + 2)
    scc = [0] * (n + 1)
    stack = [0] * (n + 2)
    v = 1
    nnodes = 0
    for node in range(1, realn + 1):
        if not mark[node]:
            continue
        nnodes += 1
        sptr[nnodes] = v
        npar[nnodes] = parent[node]
        scc[nnodes] = cc[node] + nelim[node] - 1
        height[parent[node]] = max(height[parent[node]], height[node] + 1)
        v += nvert[node]
        k = v
        shead = 1
        stack[shead] = node
        while shead > 0:
            i = stack[shead]
            shead -= 1
            k -= 1
            sperm[i] = k
            map_[i] = nnodes
            if vnext[i] != -1:
                shead += 1
                stack[shead] = vnext[i]
            if vhead[i] != -1:
                shead += 1
                stack[shead] = vhead[i]
    sptr[nnodes + 1] = v
    map_[n + 1] = nnodes + 1
    npar[nnodes + 1] = n + 1
    for i in range(realn + 1, n + 1):
        sperm[i] = i
    sparent = [0] * (nnodes + 1)
    for node in range(1, nnodes + 1):
        sparent[node] = map_[npar[node]]
    return nnodes, sptr[: nnodes + 2], sparent, scc[: nnodes + 1], sperm


def apply_perm(n, tperm, order, invp, cc):
    """core_analyse.f90:1069-1098."""
    tmp = cc[: n + 1]
    for i in range(1, n + 1):
        cc[tperm[i]] = tmp[i]
    tmp = invp[: n + 1]
    for i in range(1, n + 1):
        invp[tperm[i]] = tmp[i]
    for i in range(1, n + 1):
        order[invp[i]] = i


def find_row_lists(n, ptr, row, perm, invp, nnodes, sptr, sparent, scc):
    """core_analyse.f90:911-998."""
    seen = [0] * (n + 1)
    chead = [-1] * (nnodes + 2)
    cnext = [0] * (nnodes + 2)
    for node in range(nnodes, 0, -1):
        i = sparent[node]
        cnext[node] = chead[i]
        chead[i] = node
    rptr = [0] * (nnodes + 2)
    rlist = [0] * (sum(scc[1: nnodes + 1]) + 1)
    rptr[1] = 1
    for node in range(1, nnodes + 1):
        rptr[node + 1] = rptr[node] + scc[node]
        idx = rptr[node]
        for piv in range(sptr[node], sptr[node + 1]):
            seen[piv] = node
            rlist[idx] = piv
            idx += 1
        child = chead[node]
        while child != -1:
            for i in range(rptr[child], rptr[child + 1]):
                j = rlist[i]
                if j < sptr[node]:
                    continue
                if seen[j] == node:
                    continue
                seen[j] = node
                rlist[idx] = j
                idx += 1
            child = cnext[child]
        for piv in range(sptr[node], sptr[node + 1]):
            col = invp[piv]
            for i in range(ptr[col], ptr[col + 1]):
                j = perm[row[i]]
                if j < piv:
                    continue
                if seen[j] == node:
                    continue
                seen[j] = node
                rlist[idx] = j
                idx += 1
    return rptr, rlist


def calc_stats(nnodes, sptr, scc):
    """core_analyse.f90:862-902."""
    nfact = 0
    nflops = 0
    for node in range(1, nnodes + 1):
        nelim = sptr[node + 1] - sptr[node]
        m = scc[node] - nelim
        nfact += (nelim * (nelim + 1)) // 2
        nfact += nelim * m
        for j in range(1, nelim + 1):
            nflops += (m + j) ** 2
    return nfact, nflops


def dbl_tr_sort(n, nnodes, rptr, rlist):
    """core_analyse.f90:1007-1064: row lists into ascending order by a double transpose."""
    ptr = [0] * (n + 3)
    for node in range(1, nnodes + 1):
        for ii in range(rptr[node], rptr[node + 1]):
            ptr[rlist[ii] + 2] += 1
    ptr[1] = ptr[2] = 1
    for i in range(1, n + 1):
        ptr[i + 2] = ptr[i + 1] + ptr[i + 2]
    col = [0] * (ptr[n + 2])
    for node in range(1, nnodes + 1):
        for ii in range(rptr[node], rptr[node + 1]):
            j = rlist[ii]
            col[ptr[j + 1]] = node
            ptr[j + 1] += 1
    nptr = rptr[: nnodes + 1]
    for i in range(1, n + 1):
        for jj in range(ptr[i], ptr[i + 1]):
            node = col[jj]
            rlist[nptr[node]] = i
            nptr[node] += 1


def build_map(n, ptr, row, perm, invp, nnodes, sptr, rptr, rlist):
    """anal.f90:1129-1231: nlist(1,:) = position in val, nlist(2,:) = position in the node's lcol (both 1-based)."""
    nz = ptr[n + 1] - 1
    ptr2 = [0] * (n + 4)
    row2 = [0] * (nz + 1)
    origin = [0] * (nz + 1)
    for i in range(1, n + 1):
        for jj in range(ptr[i], ptr[i + 1]):
            k = row[jj]
            if k == i:
                continue
            ptr2[k + 2] += 1
    ptr2[1] = ptr2[2] = 1
    for i in range(1, n + 1):
        ptr2[i + 2] = ptr2[i + 2] + ptr2[i + 1]
    for i in range(1, n + 1):
        for jj in range(ptr[i], ptr[i + 1]):
            k = row[jj]
            if k == i:
                continue
            row2[ptr2[k + 1]] = i
            origin[ptr2[k + 1]] = jj
            ptr2[k + 1] += 1
    map_ = [0] * (n + 1)
    nptr = [0] * (nnodes + 2)
    nl1 = [0]
    nl2 = [0]
    pp = 1
    for node in range(1, nnodes + 1):
        blkm = rptr[node + 1] - rptr[node]
        nptr[node] = pp
        for jj in range(rptr[node], rptr[node + 1]):
            map_[rlist[jj]] = jj - rptr[node] + 1
        for j in range(sptr[node], sptr[node + 1]):
            col = invp[j]
            for i in range(ptr2[col], ptr2[col + 1]):
                k = abs(perm[row2[i]])
                if k < j:
                    continue
                nl2.append((j - sptr[node]) * blkm + map_[k])
                nl1.append(origin[i])
                pp += 1
        for j in range(sptr[node], sptr[node + 1]):
            col = invp[j]
            for ii in range(ptr[col], ptr[col + 1]):
                k = abs(perm[row[ii]])
                if k < j:
                    continue
                nl2.append((j - sptr[node]) * blkm + map_[k])
                nl1.append(ii)
                pp += 1
    nptr[nnodes + 1] = pp
    return nptr, nl1, nl2


def analyse(n, ptr0, row0, order0, nemin=32):
    """The whole analyse path on a lower-CSC pattern with 0-based ptr / row and a 0-based order (order0[i] = pivot position
    of variable i).  Returns 1-based results exactly as the Fortran holds them (lists with index 0 unused where that is
    how the Fortran indexes them): order (final pivot positions), invp, nnodes, sptr, sparent, rptr, rlist, nptr,
    nlist1, nlist2, num_factor, num_flops, flag."""
    ptr = [0] + [int(p) + 1 for p in ptr0]
    row = [0] + [int(r) + 1 for r in row0]
    order = [0] + [int(o) + 1 for o in order0]
    invp = check_order(n, order)
    if invp is None:
        return dict(flag=-8)
    aptr, arow = expand_pattern(n, ptr, row)
    # basic_analyse (core_analyse.f90:38-151)
    perm = order
    invp = [0] * (n + 1)
    for i in range(1, n + 1):
        invp[perm[i]] = i
    parent = find_etree(n, aptr, arow, perm, invp)
    realn = find_postorder(n, aptr, perm, invp, parent)
    flag = 1 if realn != n else 0
    cc = find_col_counts(n, aptr, arow, perm, invp, parent)
    nnodes, sptr, sparent, scc, tperm = find_supernodes(n, realn, parent, cc, nemin)
    apply_perm(n, tperm, perm, invp, cc)
    rptr, rlist = find_row_lists(n, aptr, arow, perm, invp, nnodes, sptr, sparent, scc)
    nfact, nflops = calc_stats(nnodes, sptr, scc)
    dbl_tr_sort(n, nnodes, rptr, rlist)
    # analyse_phase (anal.f90:997-1013)
    invp = [0] * (n + 1)
    for i in range(1, n + 1):
        invp[perm[i]] = i
    for j in range(sptr[nnodes + 1], n + 1):
        perm[invp[j]] = 0
    nptr, nl1, nl2 = build_map(n, ptr, row, perm, invp, nnodes, sptr, rptr, rlist)
    return dict(flag=flag, order=perm, invp=invp, nnodes=nnodes, sptr=sptr, sparent=sparent, rptr=rptr, rlist=rlist,
                nptr=nptr, nlist1=nl1, nlist2=nl2, num_factor=nfact, num_flops=nflops, realn=realn)
