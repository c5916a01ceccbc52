//! nalgebra's dense hot path on a B200: `gemm`, `Cholesky`, `LU`, `QR` with the reference's names, argument
//! meaning, in-place column-major storage and `Option` / `bool` error behaviour (src/base/blas.rs:729-746,
//! src/linalg/{cholesky,lu,qr}.rs).  All arithmetic runs in libnalgebra_b200 (hand-written sm_100a CUDA);
//! there is no CPU fallback -- a missing device is a panic with the library's message.
use nalgebra::{DMatrix, DVector, Dyn, PermutationSequence};
use nalgebra_b200_sys as sys;

fn check(st: i32) -> i32 {
    if st < 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::na_last_error()) }.to_string_lossy().into_owned();
        panic!("libnalgebra_b200: {msg}");
    }
    st
}

/// `c.gemm(alpha, &a, &b, beta)` (blas.rs:729-746): views with any strides are accepted; `c` is not read when
/// `beta == 0`.  Shape mismatches panic like the reference (blas_uninit.rs:244-252).
pub fn gemm(c: &mut DMatrix<f64>, alpha: f64, a: &DMatrix<f64>, b: &DMatrix<f64>, beta: f64) {
    assert_eq!(a.ncols(), b.nrows(), "gemm: dimensions mismatch for multiplication.");
    assert_eq!((c.nrows(), c.ncols()), (a.nrows(), b.ncols()), "gemm: dimensions mismatch for addition.");
    let (rsa, csa) = a.strides(); let (rsb, csb) = b.strides(); let (rsc, csc) = c.strides();
    check(unsafe {
        sys::na_dgemm(a.nrows(), a.ncols(), b.ncols(), alpha, a.as_ptr(), rsa as isize, csa as isize,
                      b.as_ptr(), rsb as isize, csb as isize, beta, c.as_mut_ptr(), rsc as isize, csc as isize)
    });
}

/// `a.tr_mul(&b)` / `gemm_tr`: the same kernel with A's strides swapped (ops.rs:674-779, blas.rs:770-803).
pub fn tr_mul(a: &DMatrix<f64>, b: &DMatrix<f64>) -> DMatrix<f64> {
    assert_eq!(a.nrows(), b.nrows(), "Matrix multiplication dimensions mismatch");
    let mut out = DMatrix::<f64>::zeros(a.ncols(), b.ncols());
    let (rsa, csa) = a.strides(); let (rsb, csb) = b.strides();
    check(unsafe {
        sys::na_dgemm(a.ncols(), a.nrows(), b.ncols(), 1.0, a.as_ptr(), csa as isize, rsa as isize,
                      b.as_ptr(), rsb as isize, csb as isize, 0.0, out.as_mut_ptr(), 1, a.ncols() as isize)
    });
    out
}

/// `Cholesky{chol}`: L in the lower triangle incl. the diagonal; the strict upper triangle keeps the caller's data.
pub struct Cholesky { chol: DMatrix<f64> }

impl Cholesky {
    /// `Cholesky::new` (cholesky.rs:196-198): `None` when a pivot is `<= 0` or NaN.
    pub fn new(matrix: DMatrix<f64>) -> Option<Self> { Self::new_internal(matrix, None) }
    /// `Cholesky::new_with_substitute` (cholesky.rs:217-219).
    pub fn new_with_substitute(matrix: DMatrix<f64>, substitute: f64) -> Option<Self> { Self::new_internal(matrix, Some(substitute)) }
    fn new_internal(mut m: DMatrix<f64>, sub: Option<f64>) -> Option<Self> {
        assert!(m.is_square(), "The input matrix must be square.");
        let n = m.nrows();
        let mut fail = 0usize;
        let st = check(unsafe { sys::na_cholesky_f64(n, m.as_mut_ptr(), n.max(1), sub.is_some() as i32, sub.unwrap_or(0.0), &mut fail) });
        (st == sys::NA_OK).then_some(Self { chol: m })
    }
    pub fn l_dirty(&self) -> &DMatrix<f64> { &self.chol }
    pub fn l(&self) -> DMatrix<f64> { self.chol.lower_triangle() }
    /// `solve_mut` (cholesky.rs:122-129).
    pub fn solve_mut(&self, b: &mut DMatrix<f64>) {
        let n = self.chol.nrows();
        assert_eq!(b.nrows(), n, "Cholesky solve matrix dimension mismatch.");
        check(unsafe { sys::na_cholesky_solve_f64(n, self.chol.as_ptr(), n.max(1), b.as_mut_ptr(), n.max(1), b.ncols()) });
    }
    pub fn solve(&self, b: &DMatrix<f64>) -> DMatrix<f64> { let mut x = b.clone(); self.solve_mut(&mut x); x }
    pub fn inverse(&self) -> DMatrix<f64> { let n = self.chol.nrows(); let mut x = DMatrix::identity(n, n); self.solve_mut(&mut x); x }
    pub fn determinant(&self) -> f64 { let p: f64 = self.chol.diagonal().iter().product(); p * p }
}

/// `LU{lu, p}`: packed factors (strict lower = L, upper = U) and the row `PermutationSequence`.
pub struct LU { lu: DMatrix<f64>, p: PermutationSequence<Dyn> }

impl LU {
    /// `LU::new` (lu.rs:93-122); never fails.
    pub fn new(mut m: DMatrix<f64>) -> Self {
        let (nr, nc) = m.shape();
        let mn = nr.min(nc);
        let mut swaps = vec![0usize; 2 * mn.max(1)];
        let mut ns = 0usize;
        check(unsafe { sys::na_lu_f64(nr, nc, m.as_mut_ptr(), nr.max(1), swaps.as_mut_ptr(), &mut ns) });
        let mut p = PermutationSequence::identity_generic(Dyn(mn));
        for s in 0..ns { p.append_permutation(swaps[2 * s], swaps[2 * s + 1]); }
        Self { lu: m, p }
    }
    pub fn lu_internal(&self) -> &DMatrix<f64> { &self.lu }
    pub fn p(&self) -> &PermutationSequence<Dyn> { &self.p }
    pub fn u(&self) -> DMatrix<f64> { let mn = self.lu.nrows().min(self.lu.ncols()); self.lu.rows(0, mn).upper_triangle() }
    pub fn l(&self) -> DMatrix<f64> {
        let mn = self.lu.nrows().min(self.lu.ncols());
        let mut l = self.lu.columns(0, mn).lower_triangle();
        l.fill_diagonal(1.0);
        l
    }
    /// `solve_mut` (lu.rs:242-260): `false` on an exactly-zero U[i,i] (b is then garbage, as in the reference).
    /// The pair list is rebuilt by replaying the sequence on an index vector (nalgebra keeps `ipiv` private).
    pub fn solve_mut(&self, b: &mut DMatrix<f64>) -> bool {
        let n = self.lu.nrows();
        assert_eq!(b.nrows(), n, "LU solve matrix dimension mismatch.");
        assert!(self.lu.is_square(), "LU solve: unable to solve a non-square system.");
        let mut idx = DVector::<f64>::from_fn(n, |i, _| i as f64);
        self.p.permute_rows(&mut idx);
        // recover the swaps (i, i2) in application order from the permuted index vector
        let mut cur: Vec<usize> = (0..n).collect();
        let mut pairs = Vec::new();
        for i in 0..n {
            let want = idx[i] as usize;
            if cur[i] != want { let j = cur.iter().position(|&v| v == want).unwrap(); cur.swap(i, j); pairs.push(i); pairs.push(j); }
        }
        let st = check(unsafe { sys::na_lu_solve_f64(n, self.lu.as_ptr(), n.max(1), pairs.as_ptr(), pairs.len() / 2, b.as_mut_ptr(), n.max(1), b.ncols()) });
        st != sys::NA_SINGULAR
    }
    pub fn solve(&self, b: &DMatrix<f64>) -> Option<DMatrix<f64>> { let mut x = b.clone(); self.solve_mut(&mut x).then_some(x) }
    pub fn try_inverse(&self) -> Option<DMatrix<f64>> { let n = self.lu.nrows(); let mut x = DMatrix::identity(n, n); self.solve_mut(&mut x).then_some(x) }
    pub fn determinant(&self) -> f64 { self.lu.diagonal().iter().product::<f64>() * self.p.determinant::<f64>() }
    pub fn is_invertible(&self) -> bool { self.lu.diagonal().iter().all(|d| *d != 0.0) }
}

/// `QR{qr, diag}` in nalgebra's storage: column i, rows i.. = unit Householder axis; strict upper = R; R[i,i] = |diag[i]|.
pub struct QR { qr: DMatrix<f64>, diag: DVector<f64> }

impl QR {
    /// `QR::new` (qr.rs:55-76); never fails.
    pub fn new(mut m: DMatrix<f64>) -> Self {
        let (nr, nc) = m.shape();
        let mut diag = DVector::zeros(nr.min(nc));
        check(unsafe { sys::na_qr_f64(nr, nc, m.as_mut_ptr(), nr.max(1), diag.as_mut_ptr()) });
        Self { qr: m, diag }
    }
    pub fn qr_internal(&self) -> &DMatrix<f64> { &self.qr }
    pub fn diag_internal(&self) -> &DVector<f64> { &self.diag }
    pub fn r(&self) -> DMatrix<f64> {
        let mn = self.diag.len();
        let mut r = self.qr.rows(0, mn).upper_triangle();
        for i in 0..mn { r[(i, i)] = self.diag[i].abs(); }
        r
    }
    /// `q()` (qr.rs:108-129).
    pub fn q(&self) -> DMatrix<f64> {
        let (nr, nc) = self.qr.shape();
        let mut q = DMatrix::zeros(nr, nr.min(nc));
        check(unsafe { sys::na_qr_q_f64(nr, nc, self.qr.as_ptr(), nr.max(1), self.diag.as_ptr(), q.as_mut_ptr(), nr.max(1)) });
        q
    }
    /// `q_tr_mul` (qr.rs:157-171).
    pub fn q_tr_mul(&self, rhs: &mut DMatrix<f64>) {
        let (nr, nc) = self.qr.shape();
        assert_eq!(rhs.nrows(), nr);
        check(unsafe { sys::na_qr_q_tr_mul_f64(nr, nc, self.qr.as_ptr(), nr.max(1), self.diag.as_ptr(), rhs.as_mut_ptr(), nr.max(1), rhs.ncols()) });
    }
    /// `solve_mut` (qr.rs:204-256): square systems; `false` when a diagonal entry of R is zero.
    pub fn solve_mut(&self, b: &mut DMatrix<f64>) -> bool {
        let n = self.qr.nrows();
        assert!(self.qr.is_square(), "QR solve: unable to solve a non-square system.");
        assert_eq!(b.nrows(), n, "QR solve matrix dimension mismatch.");
        check(unsafe { sys::na_qr_solve_f64(n, self.qr.as_ptr(), n.max(1), self.diag.as_ptr(), b.as_mut_ptr(), n.max(1), b.ncols()) }) != sys::NA_SINGULAR
    }
    pub fn is_invertible(&self) -> bool { self.diag.iter().all(|d| *d != 0.0) }
}

fn sequence(pairs: &[usize], len: usize, dim: usize) -> PermutationSequence<Dyn> {
    let mut p = PermutationSequence::identity(dim);
    for s in 0..len { p.append_permutation(pairs[2 * s], pairs[2 * s + 1]); }
    p
}

/// `FullPivLU{lu, p, q}` (full_piv_lu.rs:56-91): P * matrix * Q = L U, pivot = `icamax_full` of the trailing matrix.
/// The device applies the reference's arithmetic exactly, so `lu`, `p` and `q` are bit-identical to nalgebra's.
pub struct FullPivLU { lu: DMatrix<f64>, p: PermutationSequence<Dyn>, q: PermutationSequence<Dyn> }

impl FullPivLU {
    pub fn new(mut m: DMatrix<f64>) -> Self {
        let (nr, nc) = m.shape();
        let mn = nr.min(nc);
        let (mut ps, mut qs) = (vec![0usize; 2 * mn.max(1)], vec![0usize; 2 * mn.max(1)]);
        let (mut np, mut nq) = (0usize, 0usize);
        check(unsafe { sys::na_full_piv_lu_f64(nr, nc, m.as_mut_ptr(), nr.max(1), ps.as_mut_ptr(), &mut np, qs.as_mut_ptr(), &mut nq) });
        Self { lu: m, p: sequence(&ps, np, mn), q: sequence(&qs, nq, mn) }
    }
    pub fn lu_internal(&self) -> &DMatrix<f64> { &self.lu }
    pub fn p(&self) -> &PermutationSequence<Dyn> { &self.p }
    pub fn q(&self) -> &PermutationSequence<Dyn> { &self.q }
    pub fn u(&self) -> DMatrix<f64> { let mn = self.lu.nrows().min(self.lu.ncols()); self.lu.rows(0, mn).upper_triangle() }
    pub fn l(&self) -> DMatrix<f64> {
        let mn = self.lu.nrows().min(self.lu.ncols());
        let mut l = self.lu.columns(0, mn).into_owned();
        l.fill_upper_triangle(0.0, 1);
        l.fill_diagonal(1.0);
        l
    }
    pub fn is_invertible(&self) -> bool { let d = self.lu.nrows(); self.lu[(d - 1, d - 1)] != 0.0 }
    /// `solve_mut` (full_piv_lu.rs:189-215).
    pub fn solve_mut(&self, b: &mut DMatrix<f64>) -> bool {
        let n = self.lu.nrows();
        assert!(self.lu.is_square(), "FullPivLU solve: unable to solve a non-square system.");
        assert_eq!(b.nrows(), n, "FullPivLU solve matrix dimension mismatch.");
        if !self.is_invertible() { return false; }
        self.p.permute_rows(b);
        check(unsafe { sys::na_tri_solve_f64(1, 0, 1, n, self.lu.as_ptr(), n.max(1), b.as_mut_ptr(), n.max(1), b.ncols()) });
        check(unsafe { sys::na_tri_solve_f64(0, 0, 0, n, self.lu.as_ptr(), n.max(1), b.as_mut_ptr(), n.max(1), b.ncols()) });
        self.q.inv_permute_rows(b);
        true
    }
    pub fn determinant(&self) -> f64 {
        let d = self.lu.nrows();
        let last = self.lu[(d - 1, d - 1)];
        if last == 0.0 { return 0.0; }
        (0..d - 1).fold(last, |acc, i| acc * self.lu[(i, i)]) * self.p.determinant::<f64>() * self.q.determinant::<f64>()
    }
}

/// `ColPivQR{col_piv_qr, p, diag}` (col_piv_qr.rs:56-93): matrix * P = Q R, storage as `QR`.
pub struct ColPivQR { col_piv_qr: DMatrix<f64>, p: PermutationSequence<Dyn>, diag: DVector<f64> }

impl ColPivQR {
    pub fn new(mut m: DMatrix<f64>) -> Self {
        let (nr, nc) = m.shape();
        let mn = nr.min(nc);
        let mut diag = DVector::zeros(mn);
        let mut ps = vec![0usize; 2 * mn.max(1)];
        let mut np = 0usize;
        check(unsafe { sys::na_col_piv_qr_f64(nr, nc, m.as_mut_ptr(), nr.max(1), diag.as_mut_ptr(), ps.as_mut_ptr(), &mut np) });
        Self { col_piv_qr: m, p: sequence(&ps, np, mn), diag }
    }
    pub fn col_piv_qr_internal(&self) -> &DMatrix<f64> { &self.col_piv_qr }
    pub fn p(&self) -> &PermutationSequence<Dyn> { &self.p }
    pub fn r(&self) -> DMatrix<f64> {
        let mn = self.diag.len();
        let mut r = self.col_piv_qr.rows(0, mn).upper_triangle();
        for i in 0..mn { r[(i, i)] = self.diag[i].abs(); }
        r
    }
    /// `q()` (col_piv_qr.rs:129-150): the axes are stored like `QR`'s, so the same device routine forms Q.
    pub fn q(&self) -> DMatrix<f64> {
        let (nr, nc) = self.col_piv_qr.shape();
        let mut q = DMatrix::zeros(nr, nr.min(nc));
        check(unsafe { sys::na_qr_q_f64(nr, nc, self.col_piv_qr.as_ptr(), nr.max(1), self.diag.as_ptr(), q.as_mut_ptr(), nr.max(1)) });
        q
    }
    /// `solve_mut` (col_piv_qr.rs:227-247).
    pub fn solve_mut(&self, b: &mut DMatrix<f64>) -> bool {
        let n = self.col_piv_qr.nrows();
        assert!(self.col_piv_qr.is_square(), "ColPivQR solve: unable to solve a non-square system.");
        assert_eq!(b.nrows(), n, "ColPivQR solve matrix dimension mismatch.");
        let ok = check(unsafe { sys::na_qr_solve_f64(n, self.col_piv_qr.as_ptr(), n.max(1), self.diag.as_ptr(), b.as_mut_ptr(), n.max(1), b.ncols()) }) != sys::NA_SINGULAR;
        self.p.inv_permute_rows(b);
        ok
    }
    pub fn is_invertible(&self) -> bool { self.diag.iter().all(|d| *d != 0.0) }
    pub fn determinant(&self) -> f64 { self.diag.iter().product::<f64>() * self.p.determinant::<f64>() }
}

/// `householder::assemble_q` (householder.rs:132-152): axes in column i, rows i + 1.. of `m`.  The reflector product is
/// 1 (+) the `QR::q` of the storage one row down, so the device routine of the QR path forms it.
fn assemble_q(m: &DMatrix<f64>, signs: &DVector<f64>) -> DMatrix<f64> {
    let n = m.nrows();
    let mut q = DMatrix::zeros(n, n);
    q[(0, 0)] = 1.0;
    if n > 1 {
        check(unsafe { sys::na_qr_q_f64(n - 1, n - 1, m.as_ptr().add(1), n, signs.as_ptr(), q.as_mut_ptr().add(1 + n), n) });
    }
    q
}

/// `Hessenberg{hess, subdiag}` (hessenberg.rs:61-100): matrix = Q H Q^T.
pub struct Hessenberg { hess: DMatrix<f64>, subdiag: DVector<f64> }

impl Hessenberg {
    pub fn new(mut hess: DMatrix<f64>) -> Self {
        assert!(hess.is_square(), "Cannot compute the hessenberg decomposition of a non-square matrix.");
        let n = hess.nrows();
        assert!(n != 0, "Cannot compute the hessenberg decomposition of an empty matrix.");
        let mut subdiag = DVector::zeros(n - 1);
        check(unsafe { sys::na_hessenberg_f64(n, hess.as_mut_ptr(), n, subdiag.as_mut_ptr()) });
        Self { hess, subdiag }
    }
    pub fn hess_internal(&self) -> &DMatrix<f64> { &self.hess }
    /// `h()` (hessenberg.rs:128-140).
    pub fn h(&self) -> DMatrix<f64> {
        let n = self.hess.nrows();
        let mut res = self.hess.clone();
        res.fill_lower_triangle(0.0, 2);
        for i in 0..n - 1 { res[(i + 1, i)] = self.subdiag[i].abs(); }
        res
    }
    pub fn q(&self) -> DMatrix<f64> { assemble_q(&self.hess, &self.subdiag) }
    pub fn unpack(self) -> (DMatrix<f64>, DMatrix<f64>) { (self.q(), self.h()) }
}

/// `SymmetricTridiagonal{tri, off_diagonal}` (symmetric_tridiagonal.rs:54-95); only the lower triangle is read.
pub struct SymmetricTridiagonal { tri: DMatrix<f64>, off_diagonal: DVector<f64> }

impl SymmetricTridiagonal {
    pub fn new(mut m: DMatrix<f64>) -> Self {
        assert!(m.is_square(), "Unable to compute the symmetric tridiagonal decomposition of a non-square matrix.");
        let n = m.nrows();
        assert!(n != 0, "Unable to compute the symmetric tridiagonal decomposition of an empty matrix.");
        let mut off_diagonal = DVector::zeros(n - 1);
        check(unsafe { sys::na_symmetric_tridiagonal_f64(n, m.as_mut_ptr(), n, off_diagonal.as_mut_ptr()) });
        Self { tri: m, off_diagonal }
    }
    pub fn internal_tri(&self) -> &DMatrix<f64> { &self.tri }
    pub fn diagonal(&self) -> DVector<f64> { self.tri.diagonal() }
    pub fn off_diagonal(&self) -> DVector<f64> { self.off_diagonal.map(|e| e.abs()) }
    pub fn q(&self) -> DMatrix<f64> { assemble_q(&self.tri, &self.off_diagonal) }
    pub fn unpack(self) -> (DMatrix<f64>, DVector<f64>, DVector<f64>) { (self.q(), self.diagonal(), self.off_diagonal()) }
    pub fn unpack_tridiagonal(self) -> (DVector<f64>, DVector<f64>) { (self.diagonal(), self.off_diagonal()) }
}

/// `Bidiagonal{uv, diagonal, off_diagonal, upper_diagonal}` (bidiagonal.rs:74-150): matrix = U D V^T.
pub struct Bidiagonal { uv: DMatrix<f64>, diagonal: DVector<f64>, off_diagonal: DVector<f64>, upper_diagonal: bool }

impl Bidiagonal {
    pub fn new(mut matrix: DMatrix<f64>) -> Self {
        let (nr, nc) = matrix.shape();
        let mn = nr.min(nc);
        assert!(mn != 0, "Cannot compute the bidiagonalization of an empty matrix.");
        let mut diagonal = DVector::zeros(mn);
        let mut off_diagonal = DVector::zeros(mn - 1);
        let mut dummy = 0.0f64;
        let off_ptr = if mn > 1 { off_diagonal.as_mut_ptr() } else { &mut dummy as *mut f64 };
        check(unsafe { sys::na_bidiagonal_f64(nr, nc, matrix.as_mut_ptr(), nr, diagonal.as_mut_ptr(), off_ptr) });
        Self { uv: matrix, diagonal, off_diagonal, upper_diagonal: nr >= nc }
    }
    pub fn is_upper_diagonal(&self) -> bool { self.upper_diagonal }
    pub fn uv_internal(&self) -> &DMatrix<f64> { &self.uv }
    pub fn diagonal(&self) -> DVector<f64> { self.diagonal.map(|e| e.abs()) }
    pub fn off_diagonal(&self) -> DVector<f64> { self.off_diagonal.map(|e| e.abs()) }
    /// `d()` (bidiagonal.rs:185-200).
    pub fn d(&self) -> DMatrix<f64> {
        let mn = self.diagonal.len();
        let mut res = DMatrix::zeros(mn, mn);
        for i in 0..mn { res[(i, i)] = self.diagonal[i].abs(); }
        for i in 0..mn.saturating_sub(1) {
            if self.upper_diagonal { res[(i, i + 1)] = self.off_diagonal[i].abs(); } else { res[(i + 1, i)] = self.off_diagonal[i].abs(); }
        }
        res
    }
    /// reflector product of the axes in column i, rows i + shift.. of `st` (rows x min(rows, cols))
    fn q_of(st: &DMatrix<f64>, signs: &DVector<f64>, shift: usize) -> DMatrix<f64> {
        let (rows, cols) = st.shape();
        let k = rows.min(cols);
        let mut q = DMatrix::zeros(rows, k);
        if shift == 0 {
            check(unsafe { sys::na_qr_q_f64(rows, cols, st.as_ptr(), rows, signs.as_ptr(), q.as_mut_ptr(), rows) });
        } else {
            q[(0, 0)] = 1.0;
            if rows > 1 && k > 1 {
                check(unsafe { sys::na_qr_q_f64(rows - 1, k - 1, st.as_ptr().add(1), rows, signs.as_ptr(), q.as_mut_ptr().add(1 + rows), rows) });
            }
        }
        q
    }
    /// `u()` (bidiagonal.rs:205-240).
    pub fn u(&self) -> DMatrix<f64> {
        if self.upper_diagonal { Self::q_of(&self.uv, &self.diagonal, 0) } else { Self::q_of(&self.uv, &self.off_diagonal, 1) }
    }
    /// `v_t()` (bidiagonal.rs:244-283): the row axes are the column axes of the transposed storage.
    pub fn v_t(&self) -> DMatrix<f64> {
        let st = self.uv.transpose();
        let v = if self.upper_diagonal { Self::q_of(&st, &self.off_diagonal, 1) } else { Self::q_of(&st, &self.diagonal, 0) };
        v.transpose()
    }
    pub fn unpack(self) -> (DMatrix<f64>, DMatrix<f64>, DMatrix<f64>) { (self.u(), self.d(), self.v_t()) }
}
