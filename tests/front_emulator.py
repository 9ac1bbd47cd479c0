"""numpy emulation of the front executor's numeric phase (krets_b200/csrc/front_kernels.cuh) driven by the plan tables that
krb_analyze produced: used by the CPU-tier tests to validate the host symbolic analysis (supernodes, relative indices,
assembly lists, level order) without a GPU.  Test infrastructure only — the product has no CPU path."""
import numpy as np


def factor_and_solve(ft: dict, row_order, col_order, A_vals, b, dtype=np.float64):
    """ft: Plan.front_tables(); A_vals: stamped CSC values (entry order of Plan.pattern()); b: rhs in equation order.
    Returns x in unknown order.  Follows krb_front_factor / krb_front_forward / krb_front_backward front by front."""
    piv0, idx_ptr, idx = ft["piv0"], ft["idx_ptr"], ft["idx"]
    child_ptr, child, rel = ft["child_ptr"], ft["child"], ft["rel"]
    asm_ptr, asm_ent, asm_pos = ft["asm_ptr"], ft["asm_ent"], ft["asm_pos"]
    nf = ft["n_fronts"]
    n = len(row_order)
    bp = np.asarray(b, dtype=dtype)[np.asarray(row_order)]
    Lblk, U12, CB, CBV = [None] * nf, [None] * nf, [None] * nf, [None] * nf
    order = [f for l in range(ft["n_levels"]) for f in ft["level_fronts"][ft["level_ptr"][l]:ft["level_ptr"][l + 1]]]
    done = np.zeros(nf, dtype=bool)
    for f in order:
        p, q = piv0[f + 1] - piv0[f], idx_ptr[f + 1] - idx_ptr[f]
        m = p + q
        F = np.zeros(m * m, dtype=dtype)
        sl = slice(asm_ptr[f], asm_ptr[f + 1])
        assert len(np.unique(asm_pos[sl])) == asm_ptr[f + 1] - asm_ptr[f]
        F[asm_pos[sl]] = A_vals[asm_ent[sl]]
        F = F.reshape(m, m).T.copy()          # column-major r + m*c  ->  F[r, c]
        w = np.zeros(m, dtype=dtype)
        w[:p] = bp[piv0[f]:piv0[f + 1]]
        for c in child[child_ptr[f]:child_ptr[f + 1]]:
            assert done[c], "child scheduled after its parent"
            r = rel[idx_ptr[c]:idx_ptr[c + 1]]
            F[np.ix_(r, r)] += CB[c]
            w[r] += CBV[c]
        for k in range(p):
            if k + 1 < p:                       # threshold partial pivoting inside the fully-summed block (front_pivot_search)
                cand = np.abs(F[k + 1:p, k])
                i = k + 1 + int(np.argmax(cand))
                d = abs(F[k, k])
                if not (np.isfinite(d) and d > 0.0 and not (cand.max() > 8.0 * d)):
                    F[[k, i], :] = F[[i, k], :]
                    w[[k, i]] = w[[i, k]]
            F[k + 1:, k] /= F[k, k]
            F[k + 1:, k + 1:] -= np.outer(F[k + 1:, k], F[k, k + 1:])
        Lblk[f], U12[f], CB[f] = F[:, :p].copy(), F[:p, p:].copy(), F[p:, p:].copy()
        for k in range(p):                      # forward: unit lower
            w[k + 1:] -= F[k + 1:, k] * w[k]
        bp[piv0[f]:piv0[f + 1]] = w[:p]
        CBV[f] = w[p:].copy()
        done[f] = True
    xp = np.zeros(n, dtype=dtype)
    for f in reversed(order):
        p = piv0[f + 1] - piv0[f]
        upd = idx[idx_ptr[f]:idx_ptr[f + 1]]
        w = bp[piv0[f]:piv0[f + 1]] - U12[f] @ xp[upd]
        U11 = Lblk[f][:p, :p]
        for k in range(p - 1, -1, -1):
            w[k] /= U11[k, k]
            w[:k] -= U11[:k, k] * w[k]
        xp[piv0[f]:piv0[f + 1]] = w
    x = np.zeros(n, dtype=dtype)
    x[np.asarray(col_order)] = xp
    return x
