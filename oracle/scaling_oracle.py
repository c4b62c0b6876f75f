"""Matching-based scaling oracle: a plain-Python restatement of the reference's Hungarian scaling

    hungarian_scale_sym / hungarian_wrapper      src/spral/scaling.f90:134-170, 597-801
    hungarian_init_heurisitic / hungarian_match  src/spral/scaling.f90:810-929, 938-1194
    auction_scale_sym / auction_match / _core    src/spral/scaling.f90:269-309, 1504-1609, 1351-1489
    heap_update / heap_pop / heap_delete         src/spral/scaling.f90:1203-1325
    match_postproc (square case)                 src/spral/scaling.f90:1611-1640
    half_to_full                                 src/spral/matrix_util.f90:3167-3302
    match_order_metis / mo_match / mo_split      src/spral/match_order.f90:135-208, 495-629, 220-396 (end of this file)

(the algorithm of MC64: maximum-product matching by shortest augmenting paths on c_ij = max_i log|a_ij| - log|a_ij|; the
scaling is exp((u_i + v_i - cmax_i) / 2)).  TEST INFRASTRUCTURE ONLY: tests/ compare the product's C++ implementation
(galahad_b200/csrc/scaling.cpp) with it.  1-based lists (index 0 unused) and `math.log` / `math.exp` (the C library's, as
the C++ side) so that the comparison can be exact.  The Fortran cannot be compiled here: against SPRAL's own output the
scaling is `parity unpinned`; what is pinned are the defining properties (|s_i a_ij s_j| <= 1, = 1 on the matching).
"""
import math

RINF = 1.7976931348623157e308


def half_to_full(n, row, ptr, a):
    """matrix_util.f90:3167-3302 (rebase = 0).  row, a: lists long enough for the expanded matrix; ptr has n + 2 entries."""
    oldtau = ptr[n + 1] - 1
    iw = [0] * (n + 1)
    ndiag = 0
    for j in range(1, n + 1):
        i1 = ptr[j]
        i2 = ptr[j + 1] - 1
        iw[j] += (i2 - i1) + 1
        for ii in range(i1, i2 + 1):
            i = row[ii]
            if i != j:
                iw[i] += 1
            else:
                ndiag += 1
    newtau = 2 * oldtau - ndiag
    ipkp1 = oldtau + 1
    ckp1 = newtau + 1
    for j in range(n, 0, -1):
        i1 = ptr[j]
        i2 = ipkp1
        lenk = i2 - i1
        jstart = ckp1
        ipkp1 = i1
        i2 -= 1
        for ii in range(i2, i1 - 1, -1):
            jstart -= 1
            a[jstart] = a[ii]
            row[jstart] = row[ii]
        ptr[j] = jstart
        ckp1 -= iw[j]
        iw[j] = lenk
    for j in range(n, 0, -1):
        i1 = ptr[j]
        i2 = ptr[j] + iw[j] - 1
        for ii in range(i1, i2 + 1):
            i = row[ii]
            if i == j:
                continue
            ptr[i] -= 1
            ipos = ptr[i]
            a[ipos] = a[ii]
            row[ipos] = j
    ptr[n + 1] = newtau + 1


def heap_update(idx, q, val, l):
    pos = l[idx]
    if pos <= 1:
        q[pos] = idx
        return
    v = val[idx]
    while pos > 1:
        parent_pos = pos // 2
        parent_idx = q[parent_pos]
        if v >= val[parent_idx]:
            break
        q[pos] = parent_idx
        l[parent_idx] = pos
        pos = parent_pos
    q[pos] = idx
    l[idx] = pos


def heap_delete(pos0, qlen, q, d, l):
    """Returns the new qlen."""
    if qlen == pos0:
        return qlen - 1
    idx = q[qlen]
    v = d[idx]
    qlen -= 1
    pos = pos0
    if pos > 1:
        while True:
            parent = pos // 2
            qk = q[parent]
            if v >= d[qk]:
                break
            q[pos] = qk
            l[qk] = pos
            pos = parent
            if pos <= 1:
                break
    q[pos] = idx
    l[idx] = pos
    if pos != pos0:
        return qlen
    while True:
        child = 2 * pos
        if child > qlen:
            break
        dk = d[q[child]]
        if child < qlen:
            dr = d[q[child + 1]]
            if dk > dr:
                child += 1
                dk = dr
        if v <= dk:
            break
        qk = q[child]
        q[pos] = qk
        l[qk] = pos
        pos = child
    q[pos] = idx
    l[idx] = pos
    return qlen


def init_heuristic(m, n, ptr, row, val, iperm, jperm, dualu, d, l, search_from):
    """scaling.f90:810-929.  Returns num."""
    num = 0
    for i in range(1, m + 1):
        dualu[i] = RINF
        l[i] = 0
    for j in range(1, n + 1):
        for k in range(ptr[j], ptr[j + 1]):
            i = row[k]
            if val[k] > dualu[i]:
                continue
            dualu[i] = val[k]
            iperm[i] = j
            l[i] = k
    for i in range(1, m + 1):
        j = iperm[i]
        if j == 0:
            continue
        iperm[i] = 0
        if jperm[j] != 0:
            continue
        if (ptr[j + 1] - ptr[j] > m // 10) and (m > 50):
            continue
        num += 1
        iperm[i] = j
        jperm[j] = l[i]
    if num == min(m, n):
        return num
    for j in range(1, n + 1):
        d[j] = 0.0
        search_from[j] = ptr[j]
    for j in range(1, n + 1):
        if jperm[j] != 0:
            continue
        if ptr[j] > ptr[j + 1] - 1:
            continue
        i0 = row[ptr[j]]
        vj = val[ptr[j]] - dualu[i0]
        k0 = ptr[j]
        for k in range(ptr[j] + 1, ptr[j + 1]):
            i = row[k]
            di = val[k] - dualu[i]
            if di > vj:
                continue
            if di == vj and di != RINF:
                if iperm[i] != 0 or iperm[i0] == 0:
                    continue
            vj = di
            i0 = i
            k0 = k
        d[j] = vj
        if iperm[i0] == 0:
            num += 1
            jperm[j] = k0
            iperm[i0] = j
            search_from[j] = k0 + 1
            continue
        done = False
        for k in range(k0, ptr[j + 1]):
            i = row[k]
            if (val[k] - dualu[i]) > vj:
                continue
            jj = iperm[i]
            for kk in range(search_from[jj], ptr[jj + 1]):
                ii = row[kk]
                if iperm[ii] > 0:
                    continue
                if (val[kk] - dualu[ii]) <= d[jj]:
                    jperm[jj] = kk
                    iperm[ii] = jj
                    search_from[jj] = kk + 1
                    num += 1
                    jperm[j] = k
                    iperm[i] = j
                    search_from[j] = k + 1
                    done = True
                    break
            if done:
                break
            search_from[jj] = ptr[jj + 1]
    return num


def hungarian_match(m, n, ptr, row, val):
    """scaling.f90:938-1194.  Returns (iperm, num, dualu, dualv), 1-based lists."""
    iperm = [0] * (m + 1)
    jperm = [0] * (n + 1)
    out = [0] * (n + 1)
    pr = [0] * (n + 1)
    q = [0] * (m + 2)
    longwork = [0] * (m + 1)
    l = [0] * (m + 1)
    d = [0.0] * (max(m, n) + 1)
    dualu = [0.0] * (m + 1)
    dualv = [0.0] * (n + 1)
    num = init_heuristic(m, n, ptr, row, val, iperm, jperm, dualu, d, longwork, out)
    if num != min(m, n):
        for i in range(1, m + 1):
            d[i] = RINF
            l[i] = 0
        isp = -1
        jsp = -1
        for jord in range(1, n + 1):
            if jperm[jord] != 0:
                continue
            dmin = RINF
            qlen = 0
            low = m + 1
            up = m + 1
            csp = RINF
            j = jord
            pr[j] = -1
            for klong in range(ptr[j], ptr[j + 1]):
                i = row[klong]
                dnew = val[klong] - dualu[i]
                if dnew >= csp:
                    continue
                if iperm[i] == 0:
                    csp = dnew
                    isp = klong
                    jsp = j
                else:
                    if dnew < dmin:
                        dmin = dnew
                    d[i] = dnew
                    qlen += 1
                    longwork[qlen] = klong
            q0 = qlen
            qlen = 0
            for kk in range(1, q0 + 1):
                klong = longwork[kk]
                i = row[klong]
                if csp <= d[i]:
                    d[i] = RINF
                    continue
                if d[i] <= dmin:
                    low -= 1
                    q[low] = i
                    l[i] = low
                else:
                    qlen += 1
                    l[i] = qlen
                    heap_update(i, q, d, l)
                jj = iperm[i]
                out[jj] = klong
                pr[jj] = j
            for _ in range(1, num + 1):
                if low == up:
                    if qlen == 0:
                        break
                    i = q[1]
                    if d[i] >= csp:
                        break
                    dmin = d[i]
                    while qlen > 0:
                        i = q[1]
                        if d[i] > dmin:
                            break
                        i = q[1]
                        qlen = heap_delete(1, qlen, q, d, l)
                        low -= 1
                        q[low] = i
                        l[i] = low
                q0 = q[up - 1]
                dq0 = d[q0]
                if dq0 >= csp:
                    break
                up -= 1
                j = iperm[q0]
                vj = dq0 - val[jperm[j]] + dualu[q0]
                for klong in range(ptr[j], ptr[j + 1]):
                    i = row[klong]
                    if l[i] >= up:
                        continue
                    dnew = vj + val[klong] - dualu[i]
                    if dnew >= csp:
                        continue
                    if iperm[i] == 0:
                        csp = dnew
                        isp = klong
                        jsp = j
                    else:
                        di = d[i]
                        if di <= dnew:
                            continue
                        if l[i] >= low:
                            continue
                        d[i] = dnew
                        if dnew <= dmin:
                            lpos = l[i]
                            if lpos != 0:
                                qlen = heap_delete(lpos, qlen, q, d, l)
                            low -= 1
                            q[low] = i
                            l[i] = low
                        else:
                            if l[i] == 0:
                                qlen += 1
                                l[i] = qlen
                            heap_update(i, q, d, l)
                        jj = iperm[i]
                        out[jj] = klong
                        pr[jj] = j
            if csp != RINF:
                num += 1
                i = row[isp]
                iperm[i] = jsp
                jperm[jsp] = isp
                j = jsp
                for _ in range(1, num + 1):
                    jj = pr[j]
                    if jj == -1:
                        break
                    klong = out[j]
                    i = row[klong]
                    iperm[i] = jj
                    jperm[jj] = klong
                    j = jj
                for kk in range(up, m + 1):
                    i = q[kk]
                    dualu[i] = dualu[i] + d[i] - csp
            for kk in range(low, m + 1):
                i = q[kk]
                d[i] = RINF
                l[i] = 0
            for kk in range(1, qlen + 1):
                i = q[kk]
                d[i] = RINF
                l[i] = 0
    for j in range(1, n + 1):
        klong = jperm[j]
        dualv[j] = val[klong] - dualu[row[klong]] if klong != 0 else 0.0
    for i in range(1, m + 1):
        if iperm[i] == 0:
            dualu[i] = 0.0
    if num == min(m, n):
        return iperm, num, dualu, dualv
    jperm = [0] * (n + 1)
    k = 0
    for i in range(1, m + 1):
        if iperm[i] == 0:
            k += 1
            out[k] = i
        else:
            jperm[iperm[i]] = i
    k = 0
    for j in range(1, n + 1):
        if jperm[j] != 0:
            continue
        k += 1
        iperm[out[k]] = -j
    return iperm, num, dualu, dualv


def hungarian_scale_sym(n, ptr0, row0, val0, scale_if_singular=True):
    """hungarian_scale_sym + hungarian_wrapper(sym = .true.) on a 0-based lower-CSC matrix.
    Returns (flag, scaling (list of n), match (1-based list of n), matched)."""
    ptr = [0] + [int(p) + 1 for p in ptr0]
    row = [0] + [int(r) + 1 for r in row0]
    val = [0.0] + [float(v) for v in val0]
    flag = 0
    ne = 2 * (ptr[n + 1] - 1)
    ptr2 = [0] * (n + 2)
    row2 = [0] * (ne + 1)
    val2 = [0.0] * (ne + 1)
    klong = 1
    for i in range(1, n + 1):
        ptr2[i] = klong
        for jlong in range(ptr[i], ptr[i + 1]):
            if val[jlong] == 0.0:
                continue
            row2[klong] = row[jlong]
            val2[klong] = math.log(abs(val[jlong]))
            klong += 1
    ptr2[n + 1] = klong
    half_to_full(n, row2, ptr2, val2)
    cmax = [0.0] * (n + 1)
    for i in range(1, n + 1):
        colmax = max(val2[ptr2[i]:ptr2[i + 1]]) if ptr2[i + 1] > ptr2[i] else -RINF
        cmax[i] = colmax
        for k in range(ptr2[i], ptr2[i + 1]):
            val2[k] = colmax - val2[k]
    match, matched, dualu, dualv = hungarian_match(n, n, ptr2, row2, val2)
    rscaling = [0.0] * (n + 1)
    cscaling = [0.0] * (n + 1)
    if matched != n:
        if scale_if_singular:
            flag = 1
        else:
            flag = -2
    if matched == n:
        for i in range(1, n + 1):
            rscaling[i] = dualu[i]
            cscaling[i] = dualv[i] - cmax[i]
        # match_postproc, m == n
        rsum = 0.0               # a plain left-to-right sum (Python's sum() compensates since 3.12)
        csum = 0.0
        for i in range(1, n + 1):
            rsum += rscaling[i]
        for i in range(1, n + 1):
            csum += cscaling[i]
        ravg = rsum / n
        cavg = csum / n
        adjust = (ravg - cavg) / 2
        for i in range(1, n + 1):
            rscaling[i] -= adjust
            cscaling[i] += adjust
        scaling = [math.exp((rscaling[i] + cscaling[i]) / 2) for i in range(1, n + 1)]
        return flag, scaling, match[1:n + 1], matched
    # structurally rank deficient: the matched part alone, Duff and Pralet correction for the rest
    old_to_new = [0] * (n + 1)
    new_to_old = [0] * (n + 1)
    j = matched + 1
    k = 0
    for i in range(1, n + 1):
        if match[i] < 0:
            old_to_new[i] = -j
            j += 1
        else:
            k += 1
            old_to_new[i] = k
            new_to_old[k] = i
    ne = 0
    k = 0
    j2 = 1
    for i in range(1, n + 1):
        j1 = j2
        j2 = ptr2[i + 1]
        if match[i] < 0:
            continue
        k += 1
        if k == 1:
            ptr2[1] = 1
        for jlong in range(j1, j2):
            jj = row2[jlong]
            if match[jj] < 0:
                continue
            ne += 1
            row2[ne] = old_to_new[jj]
            val2[ne] = val2[jlong]
        ptr2[k + 1] = ne + 1
    nn = k
    if nn == 0:
        ptr2[1] = 1
    cperm, matched2, dualu, dualv = hungarian_match(nn, nn, ptr2, row2, val2)
    NEG = -RINF
    for i in range(1, n + 1):
        j = old_to_new[i]
        rscaling[i] = NEG if j < 0 else (dualu[j] + dualv[j] - cmax[i]) / 2
    match = [-1] * (n + 1)
    for i in range(1, nn + 1):
        match[new_to_old[i]] = cperm[i]
    for i in range(1, n + 1):
        if match[i] == -1:
            match[i] = old_to_new[i]
    cscale = rscaling[:]
    for i in range(1, n + 1):
        for jlong in range(ptr[i], ptr[i + 1]):
            k = row[jlong]
            if cscale[i] == NEG and cscale[k] != NEG:
                rscaling[i] = max(rscaling[i], (math.log(abs(val[jlong])) if val[jlong] != 0.0 else -math.inf) + rscaling[k])
            if cscale[k] == NEG and cscale[i] != NEG:
                rscaling[k] = max(rscaling[k], (math.log(abs(val[jlong])) if val[jlong] != 0.0 else -math.inf) + rscaling[i])
    for i in range(1, n + 1):
        if cscale[i] != NEG:
            continue
        rscaling[i] = 0.0 if rscaling[i] == NEG else -rscaling[i]
    if flag == -2:
        scaling = [1.0] * n
    else:
        scaling = [math.exp((rscaling[i] + rscaling[i]) / 2) for i in range(1, n + 1)]
    return flag, scaling, match[1:n + 1], matched2


def _f32(x):
    import struct
    return struct.unpack("f", struct.pack("f", x))[0]


def auction_match_core(m, n, ptr, row, val, dualv, max_iterations=30000, max_unchanged=(10, 100, 100),
                       min_proportion=(0.90, 0.0, 0.0), eps_initial=0.01):
    """scaling.f90:1351-1489.  dualv: in/out (1-based list).  Returns (match, dualu, iterations, unmatchable)."""
    NEGH = -RINF
    owner = [0] * (m + 1)
    nxt = list(range(n + 1))
    match = [0] * (n + 1)
    dualu = [0.0] * (m + 1)
    minmn = min(m, n)
    unmatched = minmn
    prev = -1
    nunchanged = 0
    tail = n
    eps = _f32(eps_initial)                      # a default REAL option assigned to a double
    props = [_f32(p) for p in min_proportion]
    unmatchable = 0
    itr = 1
    while itr <= max_iterations:
        if unmatched == 0:
            break
        if unmatched != prev:
            nunchanged = 0
        prev = unmatched
        nunchanged += 1
        frac = _f32(_f32(float(minmn - unmatched)) / _f32(float(minmn)))     # single-precision quotient
        if any(nunchanged >= max_unchanged[k] and frac >= props[k] for k in range(3)):
            break
        eps = min(1.0, eps + 1.0 / (n + 1))
        insert = 0
        for cptr in range(1, tail + 1):
            col = nxt[cptr]
            if match[col] != 0:
                continue
            if ptr[col] == ptr[col + 1]:
                continue
            j = ptr[col]
            bestr = row[j]
            bestu = val[j] - dualu[bestr]
            bestv = NEGH
            for j in range(ptr[col] + 1, ptr[col + 1]):
                u = val[j] - dualu[row[j]]
                if u > bestu:
                    bestv = bestu
                    bestr = row[j]
                    bestu = u
                elif u > bestv:
                    bestv = u
            if bestv == NEGH:
                bestv = 0.0
            if bestu > 0:
                dualu[bestr] = dualu[bestr] + bestu - bestv + eps
                dualv[col] = bestv - eps
                match[col] = bestr
                unmatched -= 1
                k = owner[bestr]
                owner[bestr] = col
                if k != 0:
                    match[k] = 0
                    unmatched += 1
                    insert += 1
                    nxt[insert] = k
            else:
                match[col] = -1
                unmatched -= 1
                unmatchable += 1
        tail = insert
        itr += 1
    iterations = itr - 1
    for j in range(1, n + 1):
        if match[j] == -1:
            match[j] = 0
    return match, dualu, iterations, unmatchable


def auction_scale_sym(n, ptr0, row0, val0):
    """auction_scale_sym + auction_match(expand = .true.) on a 0-based lower-CSC matrix with the reference's default
    options.  Returns (scaling, match (1-based: row i matched to column match[i], 0 unmatched), matched, iterations)."""
    ptr = [0] + [int(p) + 1 for p in ptr0]
    row = [0] + [int(r) + 1 for r in row0]
    val = [0.0] + [float(v) for v in val0]
    ne = 2 * (ptr[n + 1] - 1)
    ptr2 = [0] * (n + 2)
    row2 = [0] * (ne + 1)
    val2 = [0.0] * (ne + 1)
    klong = 1
    for i in range(1, n + 1):
        ptr2[i] = klong
        for jlong in range(ptr[i], ptr[i + 1]):
            if val[jlong] == 0.0:
                continue
            row2[klong] = row[jlong]
            val2[klong] = math.log(abs(val[jlong]))
            klong += 1
    ptr2[n + 1] = klong
    half_to_full(n, row2, ptr2, val2)
    cmax = [0.0] * (n + 1)
    for i in range(1, n + 1):
        if ptr2[i + 1] <= ptr2[i]:
            cmax[i] = 0.0
            continue
        colmax = max(val2[ptr2[i]:ptr2[i + 1]])
        cmax[i] = colmax
        for k in range(ptr2[i], ptr2[i + 1]):
            val2[k] = colmax - val2[k]
    nz2 = ptr2[n + 1] - 1
    maxentry = max(val2[1:nz2 + 1]) if nz2 > 0 else -RINF
    maxentry = 2 * maxentry + 1
    for k in range(1, nz2 + 1):
        val2[k] = maxentry - val2[k]
    cscaling = [0.0] + [-cmax[i] for i in range(1, n + 1)]
    cmatch, rscaling, iterations, unmatchable = auction_match_core(n, n, ptr2, row2, val2, cscaling)
    matched = sum(1 for j in range(1, n + 1) if cmatch[j] != 0)
    for i in range(1, n + 1):
        rscaling[i] = -rscaling[i] + maxentry
        cscaling[i] = -cscaling[i] - cmax[i]
    match = [0] * (n + 1)
    for i in range(1, n + 1):
        if cmatch[i] == 0:
            continue
        match[cmatch[i]] = i
    rsum = 0.0
    csum = 0.0
    for i in range(1, n + 1):
        rsum += rscaling[i]
    for i in range(1, n + 1):
        csum += cscaling[i]
    adjust = (rsum / n - csum / n) / 2
    scaling = [math.exp(((rscaling[i] - adjust) + (cscaling[i] + adjust)) / 2) for i in range(1, n + 1)]
    return scaling, match[1:n + 1], matched, iterations


# ----------------------------------------------------------------------------------------------------------------------
# Matching-based ordering (SSIDS options%ordering = 2)
#     match_order_metis / mo_scale / mo_match / mo_split   src/spral/match_order.f90:135-208, 406-485, 495-629, 220-396
#     expand_matrix                                         src/ssids/anal.f90:87-138
# The reference orders the compressed graph with METIS, which is not in its tree: `order_compressed` is passed in, and what
# the tests pin bit for bit is everything around it (matching, split into singletons / pairs, compressed pattern, scaling,
# expansion of the compressed order).


def expand_matrix(n, ptr0, row0, val0):
    """anal.f90:87-138.  0-based lower triangle by columns in; 1-based full matrix out (aptr has n + 2 entries)."""
    nz = int(ptr0[n]) if n > 0 else 0
    aptr = [0] * (n + 2)
    arow = [0] * (2 * nz + 2)
    aval = [0.0] * (2 * nz + 2)
    for j in range(1, n + 1):
        for kk in range(int(ptr0[j - 1]), int(ptr0[j])):
            i = int(row0[kk]) + 1
            aptr[i] += 1
            if j == i:
                continue
            aptr[j] += 1
    for j in range(2, n + 1):
        aptr[j] = aptr[j - 1] + aptr[j]
    aptr[n + 1] = aptr[n] + 1
    for j in range(1, n + 1):
        for kk in range(int(ptr0[j - 1]), int(ptr0[j])):
            i = int(row0[kk]) + 1
            atemp = float(val0[kk])
            ipos = aptr[i]
            arow[ipos] = j
            aval[ipos] = atemp
            aptr[i] = ipos - 1
            if j == i:
                continue
            jpos = aptr[j]
            arow[jpos] = i
            aval[jpos] = atemp
            aptr[j] = jpos - 1
    for j in range(1, n + 1):
        aptr[j] += 1
    return aptr, arow, aval


def mo_match(n, row2, ptr2, val2):
    """match_order.f90:495-629.  val2 = absolute values, zeros removed; the three lists are overwritten.
    Returns flag (0 / 1 = structurally singular), scale (log, 1-based), perm (1-based, -1 unmatched)."""
    cmax = [0.0] * (n + 1)
    for i in range(1, n + 1):
        colmax = 0.0
        for j in range(ptr2[i], ptr2[i + 1]):
            if val2[j] > colmax:
                colmax = val2[j]
        if colmax != 0.0:
            colmax = math.log(colmax)
        cmax[i] = colmax
    for i in range(1, n + 1):
        for j in range(ptr2[i], ptr2[i + 1]):
            val2[j] = cmax[i] - math.log(val2[j])
    cperm, rank, du, dv = hungarian_match(n, n, ptr2, row2, val2)
    scale = [0.0] * (n + 1)
    if rank == n:
        for i in range(1, n + 1):
            scale[i] = (du[i] + dv[i] - cmax[i]) / 2
        return 0, scale, list(cperm)
    old_to_new = [0] * (n + 1)
    new_to_old = [0] * (n + 1)
    k = 0
    for i in range(1, n + 1):
        if cperm[i] < 0:
            old_to_new[i] = -1
        else:
            k += 1
            old_to_new[i] = k
            new_to_old[k] = i
    nne = 0
    k = 0
    ptr2[1] = 1
    j2 = 1
    for i in range(1, n + 1):
        j1 = j2
        j2 = ptr2[i + 1]
        if cperm[i] < 0:
            continue
        k += 1
        for jl in range(j1, j2):
            jj = row2[jl]
            if cperm[jj] < 0:
                continue
            nne += 1
            row2[nne] = old_to_new[jj]
            val2[nne] = val2[jl]
        ptr2[k + 1] = nne + 1
    nn = k
    cperm2, rank2, du, dv = hungarian_match(nn, nn, ptr2, row2, val2)
    for i in range(1, n + 1):
        j = old_to_new[i]
        scale[i] = -RINF if j < 0 else (du[j] + dv[j] - cmax[i]) / 2
    perm = [-1] * (n + 1)
    for i in range(1, nn + 1):
        perm[new_to_old[i]] = new_to_old[cperm2[i]]
    return 1, scale, perm


def mo_split(n, row2, ptr2, cperm, order_compressed):
    """match_order.f90:220-396.  Returns order (1-based positions, list of n + 1), the split matching, (ncomp, ptr3, row3)."""
    iwork = [0] * (n + 1)
    for i in range(1, n + 1):
        if iwork[i] != 0:
            continue
        j = i
        while True:
            if cperm[j] == -1:
                iwork[j] = -2
                break
            elif cperm[j] == i:
                iwork[j] = -1
                break
            jj = cperm[j]
            iwork[j] = jj
            iwork[jj] = j
            j = cperm[jj]
            if j == i:
                break
    cperm = list(iwork)
    old_to_new = [0] * (n + 1)
    new_to_old = [0] * (n + 1)
    k = 1
    for i in range(1, n + 1):
        j = cperm[i]
        if j < i and j > 0:
            continue
        old_to_new[i] = k
        new_to_old[k] = i
        if j > 0:
            old_to_new[j] = k
        k += 1
    ncomp_matched = k - 1
    ne = ptr2[n + 1] - 1
    ptr3 = [0] * (n + 2)
    row3 = [0] * (ne + 2)
    iwork = [0] * (n + 1)
    ptr3[1] = 1
    ncomp = 1
    jj = 1
    for i in range(1, n + 1):
        j = cperm[i]
        if j < i and j > 0:
            continue
        for kl in range(ptr2[i], ptr2[i + 1]):
            krow = old_to_new[row2[kl]]
            if iwork[krow] == i:
                continue
            if krow > ncomp_matched:
                continue
            row3[jj] = krow
            jj += 1
            iwork[krow] = i
        if j > 0:
            for kl in range(ptr2[j], ptr2[j + 1]):
                krow = old_to_new[row2[kl]]
                if iwork[krow] == i:
                    continue
                if krow > ncomp_matched:
                    continue
                row3[jj] = krow
                jj += 1
                iwork[krow] = i
        ptr3[ncomp + 1] = jj
        ncomp += 1
    ncomp -= 1
    ptr3[1] = 1
    jj = 1
    j1 = 1
    for i in range(1, ncomp + 1):
        j2 = ptr3[i + 1]
        for k in range(j1, j2):
            krow = row3[k]
            if krow < i:
                continue
            row3[jj] = krow
            jj += 1
        ptr3[i + 1] = jj
        j1 = j2
    corder = order_compressed(ncomp, ptr3, row3)          # corder[i] = position of compressed vertex i, both 1-based
    iw = [0] * (ncomp + 1)
    for i in range(1, ncomp + 1):
        iw[corder[i]] = i
    order = [0] * (n + 1)
    k = 1
    for i in range(1, ncomp + 1):
        j = new_to_old[iw[i]]
        order[j] = k
        k += 1
        if cperm[j] > 0:
            j = cperm[j]
            order[j] = k
            k += 1
    return order, cperm, (ncomp, ptr3, row3)


def match_order(n, ptr0, row0, val0, order_compressed):
    """match_order_metis (match_order.f90:135-208) on the full matrix expand_matrix makes of the 0-based lower triangle.
    Returns flag, order (0-based positions, n), scale (n), pairs (n; 1-based partner, -1 singleton, -2 unmatched) and the
    compressed lower-triangular pattern (ncomp, ptr3, row3; 1-based lists)."""
    aptr, arow, aval = expand_matrix(n, ptr0, row0, val0)
    ptr2 = [0] * (n + 2)
    row2 = [0] * len(arow)
    val2 = [0.0] * len(aval)
    k = 1
    for i in range(1, n + 1):
        ptr2[i] = k
        for j in range(aptr[i], aptr[i + 1]):
            if aval[j] == 0.0:
                continue
            row2[k] = arow[j]
            val2[k] = abs(aval[j])
            k += 1
    ptr2[n + 1] = k
    flag, scale, cperm = mo_match(n, list(row2), list(ptr2), list(val2))      # mo_scale works on its own copy
    order, pairs, comp = mo_split(n, row2, ptr2, cperm, order_compressed)
    sc = [math.exp(scale[i]) if scale[i] != -RINF else 0.0 for i in range(1, n + 1)]
    return flag, [order[i] - 1 for i in range(1, n + 1)], sc, pairs[1:n + 1], comp
