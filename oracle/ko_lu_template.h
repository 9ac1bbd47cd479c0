/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product path.
 *
 * Scalar-generic left-looking sparse LU with row partial pivoting (Gilbert–Peierls), the
 * published algorithm behind `faer::sparse::SparseColMat::sp_lu()` + `Solve::solve`, which
 * the reference calls at
 *     crates/krets-solver/src/solver/op.rs:48-58, dc.rs:81-90, ac.rs:67-95, transient.rs:74-83.
 * faer 0.23.2 (Cargo.toml:27, Cargo.lock:1976-1979) is NOT vendored under /root/reference, so
 * this is a restatement of the algorithm, not of faer's source: parity of the LU itself is
 * UNPINNED at the bit level (no reference test pins pivot order or rounding); results are
 * pinned only through the reference's own test tolerances and the AC golden parquet.
 *
 * Included twice by krets_oracle.c with
 *     KO_T      scalar type            (double | ko_c64)
 *     KO_SFX    function-name suffix   (_r | _c)
 * and the scalar ops KO_ZERO, KO_ISZERO, KO_ABS2, KO_SUB_MUL(x,a,b) (x -= a*b), KO_DIV(a,b).
 *
 * Column order: natural (faer uses COLAMD; the order only changes rounding, see above).
 */

#define KO_CAT_(a, b) a##b
#define KO_CAT(a, b) KO_CAT_(a, b)
#define KO_FN(name) KO_CAT(name, KO_SFX)

typedef struct {
    int n;
    int *Lp, *Li; KO_T *Lx;   /* unit lower, CSC, pivot (1.0) stored first in each column; row ids are ORIGINAL rows until finalised */
    int *Up, *Ui; KO_T *Ux;   /* upper, CSC, diagonal last; row ids are pivot positions */
    int *pinv;                /* pinv[orig_row] = pivot position */
    int lnz, unz, lcap, ucap;
} KO_FN(ko_lu);

static void KO_FN(ko_lu_free)(KO_FN(ko_lu) * f) {
    free(f->Lp); free(f->Li); free(f->Lx); free(f->Up); free(f->Ui); free(f->Ux); free(f->pinv);
    memset(f, 0, sizeof(*f));
}

/* depth-first search of the graph of L from node j (CSparse cs_dfs shape) */
static int KO_FN(ko_dfs)(int j, const int *Lp, const int *Li, int top, int *xi, int *pstack, const int *pinv, int *marked, int stamp) {
    int head = 0;
    xi[0] = j;
    while (head >= 0) {
        j = xi[head];
        int jnew = pinv[j];
        if (marked[j] != stamp) {
            marked[j] = stamp;
            pstack[head] = (jnew < 0) ? 0 : Lp[jnew];
        }
        int done = 1;
        int p2 = (jnew < 0) ? 0 : Lp[jnew + 1];
        for (int p = pstack[head]; p < p2; p++) {
            int i = Li[p];
            if (marked[i] == stamp) continue;
            pstack[head] = p;
            xi[++head] = i;
            done = 0;
            break;
        }
        if (done) {
            head--;
            xi[--top] = j;
        }
    }
    return top;
}

/*
 * Factor A (CSC, n x n, sorted or unsorted rows, explicit zeros kept as the reference keeps them).
 * Returns 0 on success, 1 if a column has no usable pivot (structurally or numerically singular) —
 * the condition the reference maps to Error::DecompositionFailed / MatrixDecomposition.
 */
static int KO_FN(ko_lu_factor)(int n, const int *Ap, const int *Ai, const KO_T *Ax, KO_FN(ko_lu) * f) {
    memset(f, 0, sizeof(*f));
    f->n = n;
    int annz = Ap[n];
    f->lcap = 4 * annz + n + 16;
    f->ucap = 4 * annz + n + 16;
    f->Lp = (int *)malloc(sizeof(int) * (n + 1));
    f->Up = (int *)malloc(sizeof(int) * (n + 1));
    f->Li = (int *)malloc(sizeof(int) * f->lcap);
    f->Ui = (int *)malloc(sizeof(int) * f->ucap);
    f->Lx = (KO_T *)malloc(sizeof(KO_T) * f->lcap);
    f->Ux = (KO_T *)malloc(sizeof(KO_T) * f->ucap);
    f->pinv = (int *)malloc(sizeof(int) * n);
    KO_T *x = (KO_T *)malloc(sizeof(KO_T) * n);
    int *xi = (int *)malloc(sizeof(int) * 2 * n);
    int *marked = (int *)calloc(n, sizeof(int));
    int rc = 0;
    for (int i = 0; i < n; i++) { x[i] = KO_ZERO; f->pinv[i] = -1; }
    int lnz = 0, unz = 0;
    for (int k = 0; k < n; k++) {
        f->Lp[k] = lnz;
        f->Up[k] = unz;
        if (lnz + n > f->lcap) {
            f->lcap = 2 * f->lcap + n;
            f->Li = (int *)realloc(f->Li, sizeof(int) * f->lcap);
            f->Lx = (KO_T *)realloc(f->Lx, sizeof(KO_T) * f->lcap);
        }
        if (unz + n > f->ucap) {
            f->ucap = 2 * f->ucap + n;
            f->Ui = (int *)realloc(f->Ui, sizeof(int) * f->ucap);
            f->Ux = (KO_T *)realloc(f->Ux, sizeof(KO_T) * f->ucap);
        }
        /* reach of column k of A in the graph of L -> topological order in xi[top..n) */
        int top = n;
        int stamp = k + 1;
        for (int p = Ap[k]; p < Ap[k + 1]; p++)
            if (marked[Ai[p]] != stamp) top = KO_FN(ko_dfs)(Ai[p], f->Lp, f->Li, top, xi, xi + n, f->pinv, marked, stamp);
        for (int p = top; p < n; p++) x[xi[p]] = KO_ZERO;
        for (int p = Ap[k]; p < Ap[k + 1]; p++) x[Ai[p]] = Ax[p];   /* CSC has unique (row,col) after sum_triplets */
        /* sparse triangular solve x = L \ A(:,k) */
        for (int px = top; px < n; px++) {
            int j = xi[px];
            int J = f->pinv[j];
            if (J < 0) continue;                    /* row j not yet pivotal: stays in the L part */
            /* L has unit diagonal: x[j] is final; eliminate below */
            for (int p = f->Lp[J] + 1; p < f->Lp[J + 1]; p++) KO_SUB_MUL(x[f->Li[p]], f->Lx[p], x[j]);
        }
        /* pick the pivot among non-pivotal rows: largest magnitude (plain partial pivoting) */
        int ipiv = -1;
        double amax = -1.0;
        for (int p = top; p < n; p++) {
            int i = xi[p];
            if (f->pinv[i] < 0) {
                double t = KO_ABS2(x[i]);
                if (t > amax) { amax = t; ipiv = i; }
            } else {
                f->Ui[unz] = f->pinv[i];
                f->Ux[unz++] = x[i];
            }
        }
        if (ipiv < 0 || !(amax > 0.0)) { rc = 1; break; }
        KO_T pivot = x[ipiv];
        f->Ui[unz] = k;
        f->Ux[unz++] = pivot;
        f->pinv[ipiv] = k;
        f->Li[lnz] = ipiv;
        f->Lx[lnz++] = KO_ONE;
        for (int p = top; p < n; p++) {
            int i = xi[p];
            if (f->pinv[i] < 0) {
                f->Li[lnz] = i;
                f->Lx[lnz++] = KO_DIV(x[i], pivot);
            }
            x[i] = KO_ZERO;
        }
    }
    if (!rc) {
        f->Lp[n] = lnz;
        f->Up[n] = unz;
        for (int p = 0; p < lnz; p++) f->Li[p] = f->pinv[f->Li[p]];   /* row ids -> pivot positions */
    }
    f->lnz = lnz; f->unz = unz;
    free(x); free(xi); free(marked);
    if (rc) KO_FN(ko_lu_free)(f);
    return rc;
}

/* x = A^{-1} b  (b, x dense length n; b is not modified) */
static void KO_FN(ko_lu_solve)(const KO_FN(ko_lu) * f, const KO_T *b, KO_T *x) {
    int n = f->n;
    KO_T *y = (KO_T *)malloc(sizeof(KO_T) * n);
    for (int i = 0; i < n; i++) y[f->pinv[i]] = b[i];                 /* y = P b */
    for (int j = 0; j < n; j++)                                        /* L y = y (unit diag first) */
        for (int p = f->Lp[j] + 1; p < f->Lp[j + 1]; p++) KO_SUB_MUL(y[f->Li[p]], f->Lx[p], y[j]);
    for (int j = n - 1; j >= 0; j--) {                                 /* U x = y (diag last) */
        y[j] = KO_DIV(y[j], f->Ux[f->Up[j + 1] - 1]);
        for (int p = f->Up[j]; p < f->Up[j + 1] - 1; p++) KO_SUB_MUL(y[f->Ui[p]], f->Ux[p], y[j]);
    }
    for (int i = 0; i < n; i++) x[i] = y[i];                           /* natural column order: Q = I */
    free(y);
}

#undef KO_FN
#undef KO_CAT
#undef KO_CAT_
