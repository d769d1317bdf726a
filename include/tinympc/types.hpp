// Host-side data model of the B200-native TinyMPC: the reference's four structs with every field name kept
// (/root/reference/src/tinympc/types.hpp:26-107), on an Eigen-free column-major matrix type.  Sizes are
// run-time values (the reference fixes them with the NSTATES/NINPUTS/NHORIZON macros of glob_opts.hpp:3-9);
// the scalar type is `float` (what tiny_codegen emits, codegen.cpp:152) unless TINYMPC_DOUBLE is defined
// (the shipped glob_opts.hpp:3).
#pragma once
#include <cstddef>
#include <vector>

#ifdef TINYMPC_DOUBLE
typedef double tinytype;
#else
typedef float tinytype;
#endif

// Minimal dense column-major matrix: element (i, j) at data()[i + j*rows()] -- the memory layout of the
// reference's Eigen::Matrix<tinytype, R, C> members, so a "nx x N" trajectory is [stage][state] contiguous.
class tiny_Matrix {
public:
    tiny_Matrix() : r_(0), c_(0) {}
    tiny_Matrix(int rows, int cols) : r_(rows), c_(cols), a_((size_t)rows * cols, tinytype(0)) {}
    void resize(int rows, int cols) { r_ = rows; c_ = cols; a_.assign((size_t)rows * cols, tinytype(0)); }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    tinytype *data() { return a_.data(); }
    const tinytype *data() const { return a_.data(); }
    tinytype &operator()(int i, int j) { return a_[(size_t)i + (size_t)j * r_]; }
    tinytype operator()(int i, int j) const { return a_[(size_t)i + (size_t)j * r_]; }
    tinytype &operator()(int i) { return a_[i]; }
    tinytype operator()(int i) const { return a_[i]; }
    tinytype *col(int j) { return a_.data() + (size_t)j * r_; }
    const tinytype *col(int j) const { return a_.data() + (size_t)j * r_; }
    void setZero() { a_.assign(a_.size(), tinytype(0)); }
    void setConstant(tinytype v) { a_.assign(a_.size(), v); }
    void setCol(int j, const tinytype *src) { for (int i = 0; i < r_; ++i) (*this)(i, j) = src[i]; }
private:
    int r_, c_;
    std::vector<tinytype> a_;
};

typedef tiny_Matrix tiny_VectorNx, tiny_VectorNu, tiny_MatrixNxNx, tiny_MatrixNxNu, tiny_MatrixNuNx, tiny_MatrixNuNu,
    tiny_MatrixNxNh, tiny_MatrixNuNhm1;

/** Matrices that must be recomputed with changes in time step, rho (types.hpp:26-34) */
typedef struct {
    tinytype rho;
    tiny_MatrixNuNx Kinf;
    tiny_MatrixNxNx Pinf;
    tiny_MatrixNuNu Quu_inv;
    tiny_MatrixNxNx AmBKt;
    tiny_MatrixNxNu coeff_d2p;  // carried for completeness; unused by the solver (admm.cpp:20)
} TinyCache;

/** User settings (types.hpp:39-47) */
typedef struct {
    tinytype abs_pri_tol;
    tinytype abs_dua_tol;
    int max_iter;
    int check_termination;
    int en_state_bound;
    int en_input_bound;
} TinySettings;

/** Problem variables (types.hpp:52-97) */
typedef struct {
    tiny_MatrixNxNh x;      // state trajectory, nx x N
    tiny_MatrixNuNhm1 u;    // input trajectory, nu x (N-1)
    tiny_MatrixNxNh q;      // linear cost terms
    tiny_MatrixNuNhm1 r;
    tiny_MatrixNxNh p;      // Riccati backward pass terms
    tiny_MatrixNuNhm1 d;
    tiny_MatrixNxNh v;      // slack variables
    tiny_MatrixNxNh vnew;
    tiny_MatrixNuNhm1 z;
    tiny_MatrixNuNhm1 znew;
    tiny_MatrixNxNh g;      // duals
    tiny_MatrixNuNhm1 y;

    tinytype primal_residual_state;
    tinytype primal_residual_input;
    tinytype dual_residual_state;
    tinytype dual_residual_input;
    int status;
    int iter;

    tiny_VectorNx Q;
    tiny_VectorNu R;
    tiny_MatrixNxNx Adyn;
    tiny_MatrixNxNu Bdyn;

    tiny_MatrixNuNhm1 u_min;
    tiny_MatrixNuNhm1 u_max;
    tiny_MatrixNxNh x_min;
    tiny_MatrixNxNh x_max;
    tiny_MatrixNxNh Xref;
    tiny_MatrixNuNhm1 Uref;  // ignored by the solver, as in the reference (admm.cpp:79)

    tiny_VectorNu Qu;
} TinyWorkspace;

/** Main solver structure (types.hpp:102-107) + the run-time sizes and the device backend handle */
typedef struct {
    TinySettings *settings;
    TinyCache *cache;
    TinyWorkspace *work;
    int nx, nu, N;   // replaces NSTATES / NINPUTS / NHORIZON
    void *backend;   // tmpc_ctx* (include/tmpc.h), created lazily by the first solve; owned by tiny_free
} TinySolver;
