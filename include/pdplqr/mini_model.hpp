// Minimal stand-ins for the reference's data contract, used ONLY when Eigen3 is not available
// (this image has no Eigen).  Field names, layouts and constructor arguments follow
//   /root/reference include/clqr/typedefs.hpp:8-24   (scalar, VectorXs, MatrixXs: FP64, column-major)
//   /root/reference include/clqr/lqr_model.hpp:8-89  (Node, LQRModel)
// so that code written against them compiles unchanged against the real Eigen-based headers.
#pragma once
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <vector>

namespace lqr {

using scalar = double;

class VectorXs {
public:
    VectorXs() = default;
    explicit VectorXs(int n) : v_(n, 0.0) {}
    void resize(int n) { v_.assign(n, 0.0); }
    void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }
    void setConstant(scalar a) { std::fill(v_.begin(), v_.end(), a); }
    int size() const { return (int)v_.size(); }
    scalar* data() { return v_.data(); }
    const scalar* data() const { return v_.data(); }
    scalar& operator()(int i) { return v_[i]; }
    scalar operator()(int i) const { return v_[i]; }
    scalar& operator[](int i) { return v_[i]; }
    scalar operator[](int i) const { return v_[i]; }

private:
    std::vector<scalar> v_;
};

class MatrixXs {  // column-major, like Eigen::MatrixXd
public:
    MatrixXs() = default;
    MatrixXs(int r, int c) : r_(r), c_(c), v_((size_t)r * c, 0.0) {}
    void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, 0.0); }
    void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }
    void setIdentity() { setZero(); for (int i = 0; i < (r_ < c_ ? r_ : c_); ++i) (*this)(i, i) = 1.0; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    scalar* data() { return v_.data(); }
    const scalar* data() const { return v_.data(); }
    scalar& operator()(int i, int j) { return v_[i + (size_t)j * r_]; }
    scalar operator()(int i, int j) const { return v_[i + (size_t)j * r_]; }

private:
    int r_ = 0, c_ = 0;
    std::vector<scalar> v_;
};

constexpr scalar LQR_INFTY = std::numeric_limits<scalar>::infinity();

struct Node {  // lqr_model.hpp:8-64
    int n, m, n_con;
    MatrixXs E;  // [B A]
    VectorXs c;
    MatrixXs H;  // [R S; S^T Q]
    VectorXs h;  // [r; q]
    MatrixXs D_con;  // [Du Dx]
    VectorXs e_lb, e_ub;
    bool is_terminal;
    int time_step;
    Node(int state_dim, int control_dim, int n_constraints, int time_step_, bool is_terminal_stage = false)
        : n(state_dim), m(control_dim), n_con(n_constraints), is_terminal(is_terminal_stage), time_step(time_step_) {
        if (is_terminal) { H.resize(n, n); h.resize(n); }
        else { E.resize(n, n + m); c.resize(n); H.resize(n + m, n + m); h.resize(n + m); }
        if (n_con > 0) { D_con.resize(n_con, is_terminal ? n : n + m); e_lb.resize(n_con); e_ub.resize(n_con); }
    }
    int get_constraint_dim() const { return n_con; }
};

struct LQRModel {  // lqr_model.hpp:66-89
    int n, m, N;
    std::vector<int> ncs;
    std::vector<Node> nodes;
    LQRModel(int n_, int m_, int horizon) : n(n_), m(m_), N(horizon) {
        if (N < 1) throw std::runtime_error("Horizon must be at least 1.");
        ncs.resize(N + 1);
        nodes.reserve(N + 1);
    }
    Node& get_node(int k) { return nodes[k]; }
    const Node& get_node(int k) const { return nodes[k]; }
    void add_node(int n_, int m_, int nc, int time_step, bool is_terminal_stage = false) {
        nodes.emplace_back(n_, m_, nc, time_step, is_terminal_stage);
        ncs[time_step] = nc;
    }
};

}  // namespace lqr
