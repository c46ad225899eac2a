// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity status: see oracle/README.md (pinned to the reference's own sources via oracle/_ref).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything under oracle/.  The product path (auvsl_neural_ode_b200/csrc) never includes it.
//
// Scalar layer of the CPU restatement.  The reference runs all math on CppAD::AD<double>
// (/root/reference/src/types/Scalars.h:11) and gets d(loss)/d(params) from
// CppAD::ADFun::Reverse(1) (/root/reference/src/Trainer.cpp:607,650-657).  CppAD is not in this
// image, so this file provides the two scalar types the templated restatement is instantiated on:
//   * double              – values (CppAD with no tape recording is plain double arithmetic)
//   * orc::Rev            – a minimal reverse-mode tape with CppAD's conventions:
//       CondExpGt/Lt/Eq select a branch by VALUE and differentiate the selected branch only,
//       abs'(0) = 0 (CppAD sign(0)=0), pow(x,y) differentiates both arguments.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc {

// ------------------------------------------------------------------ double overloads
inline double value(double x) { return x; }
inline double cond_gt(double l, double r, double a, double b) { return l > r ? a : b; }   // CppAD::CondExpGt
inline double cond_lt(double l, double r, double a, double b) { return l < r ? a : b; }   // CppAD::CondExpLt
inline double cond_eq(double l, double r, double a, double b) { return l == r ? a : b; }  // CppAD::CondExpEq
inline double s_abs(double x) { return std::fabs(x); }
inline double s_sqrt(double x) { return std::sqrt(x); }
inline double s_sin(double x) { return std::sin(x); }
inline double s_cos(double x) { return std::cos(x); }
inline double s_tan(double x) { return std::tan(x); }
inline double s_tanh(double x) { return std::tanh(x); }
inline double s_exp(double x) { return std::exp(x); }
inline double s_acos(double x) { return std::acos(x); }
inline double s_atan2(double y, double x) { return std::atan2(y, x); }
inline double s_pow(double x, double y) { return std::pow(x, y); }

// ------------------------------------------------------------------ reverse tape
struct Node {
    int32_t a, b;   // parent node ids (-1 = none)
    double da, db;  // local partials
};

struct Tape {
    std::vector<Node> nodes;
    void clear() { nodes.clear(); }
    int push(int a, double da, int b, double db) {
        nodes.push_back(Node{a, b, da, db});
        return (int)nodes.size() - 1;
    }
    // adj must be sized nodes.size() and hold the seeds; on return it holds all adjoints.
    void reverse(std::vector<double>& adj) const {
        for (int i = (int)nodes.size() - 1; i >= 0; --i) {
            const double w = adj[i];
            if (w == 0.0) continue;
            const Node& n = nodes[i];
            if (n.a >= 0) adj[n.a] += w * n.da;
            if (n.b >= 0) adj[n.b] += w * n.db;
        }
    }
};

inline Tape*& active_tape() {
    static thread_local Tape* t = nullptr;
    return t;
}

struct Rev {
    double v;
    int id;  // -1: constant / parameter-free value
    Rev() : v(0.0), id(-1) {}
    Rev(double x) : v(x), id(-1) {}
    Rev(int x) : v((double)x), id(-1) {}
    Rev(double x, int i) : v(x), id(i) {}
    static Rev independent(double x) { return Rev(x, active_tape()->push(-1, 0.0, -1, 0.0)); }
};

inline Rev mk1(double v, const Rev& a, double da) {
    if (a.id < 0) return Rev(v);
    return Rev(v, active_tape()->push(a.id, da, -1, 0.0));
}
inline Rev mk2(double v, const Rev& a, double da, const Rev& b, double db) {
    if (a.id < 0 && b.id < 0) return Rev(v);
    return Rev(v, active_tape()->push(a.id, da, b.id, db));
}

inline Rev operator+(const Rev& a, const Rev& b) { return mk2(a.v + b.v, a, 1.0, b, 1.0); }
inline Rev operator-(const Rev& a, const Rev& b) { return mk2(a.v - b.v, a, 1.0, b, -1.0); }
inline Rev operator*(const Rev& a, const Rev& b) { return mk2(a.v * b.v, a, b.v, b, a.v); }
inline Rev operator/(const Rev& a, const Rev& b) {
    const double q = a.v / b.v;
    return mk2(q, a, 1.0 / b.v, b, -q / b.v);
}
inline Rev operator-(const Rev& a) { return mk1(-a.v, a, -1.0); }
inline Rev& operator+=(Rev& a, const Rev& b) { a = a + b; return a; }
inline Rev& operator-=(Rev& a, const Rev& b) { a = a - b; return a; }
inline Rev& operator*=(Rev& a, const Rev& b) { a = a * b; return a; }
inline Rev& operator/=(Rev& a, const Rev& b) { a = a / b; return a; }
inline bool operator<(const Rev& a, const Rev& b) { return a.v < b.v; }
inline bool operator>(const Rev& a, const Rev& b) { return a.v > b.v; }
inline bool operator==(const Rev& a, const Rev& b) { return a.v == b.v; }

inline double value(const Rev& x) { return x.v; }
inline Rev cond_gt(const Rev& l, const Rev& r, const Rev& a, const Rev& b) { return l.v > r.v ? a : b; }
inline Rev cond_lt(const Rev& l, const Rev& r, const Rev& a, const Rev& b) { return l.v < r.v ? a : b; }
inline Rev cond_eq(const Rev& l, const Rev& r, const Rev& a, const Rev& b) { return l.v == r.v ? a : b; }
inline Rev s_abs(const Rev& x) { return mk1(std::fabs(x.v), x, x.v > 0 ? 1.0 : (x.v < 0 ? -1.0 : 0.0)); }
inline Rev s_sqrt(const Rev& x) { const double r = std::sqrt(x.v); return mk1(r, x, 0.5 / r); }
inline Rev s_sin(const Rev& x) { return mk1(std::sin(x.v), x, std::cos(x.v)); }
inline Rev s_cos(const Rev& x) { return mk1(std::cos(x.v), x, -std::sin(x.v)); }
inline Rev s_tan(const Rev& x) { const double t = std::tan(x.v); return mk1(t, x, 1.0 + t * t); }
inline Rev s_tanh(const Rev& x) { const double t = std::tanh(x.v); return mk1(t, x, 1.0 - t * t); }
inline Rev s_exp(const Rev& x) { const double e = std::exp(x.v); return mk1(e, x, e); }
inline Rev s_acos(const Rev& x) { return mk1(std::acos(x.v), x, -1.0 / std::sqrt(1.0 - x.v * x.v)); }
inline Rev s_atan2(const Rev& y, const Rev& x) {
    const double d = x.v * x.v + y.v * y.v;
    return mk2(std::atan2(y.v, x.v), y, x.v / d, x, -y.v / d);
}
inline Rev s_pow(const Rev& x, const Rev& y) {
    const double p = std::pow(x.v, y.v);
    const double dx = (x.v == 0.0) ? 0.0 : y.v * p / x.v;
    const double dy = (x.v > 0.0) ? p * std::log(x.v) : 0.0;
    return mk2(p, x, dx, y, dy);
}

// ------------------------------------------------------------------ forward-mode dual (independent gradient check)
template <int NP>
struct Dual {
    double v;
    double d[NP];
    Dual() : v(0) { for (int i = 0; i < NP; i++) d[i] = 0; }
    Dual(double x) : v(x) { for (int i = 0; i < NP; i++) d[i] = 0; }
    Dual(int x) : v(x) { for (int i = 0; i < NP; i++) d[i] = 0; }
};
#define ORC_DUAL_UN(name, val, der)                                    \
    template <int NP> inline Dual<NP> name(const Dual<NP>& x) {        \
        Dual<NP> r; r.v = (val); const double g = (der);               \
        for (int i = 0; i < NP; i++) r.d[i] = g * x.d[i]; return r; }
template <int NP> inline Dual<NP> operator+(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v + b.v; for (int i = 0; i < NP; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v - b.v; for (int i = 0; i < NP; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NP> inline Dual<NP> operator*(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v * b.v; for (int i = 0; i < NP; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NP> inline Dual<NP> operator/(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v / b.v; for (int i = 0; i < NP; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a) { Dual<NP> r; r.v = -a.v; for (int i = 0; i < NP; i++) r.d[i] = -a.d[i]; return r; }
template <int NP> inline Dual<NP> operator+(double a, const Dual<NP>& b) { return Dual<NP>(a) + b; }
template <int NP> inline Dual<NP> operator-(double a, const Dual<NP>& b) { return Dual<NP>(a) - b; }
template <int NP> inline Dual<NP> operator*(double a, const Dual<NP>& b) { return Dual<NP>(a) * b; }
template <int NP> inline Dual<NP> operator/(double a, const Dual<NP>& b) { return Dual<NP>(a) / b; }
template <int NP> inline Dual<NP> operator+(const Dual<NP>& a, double b) { return a + Dual<NP>(b); }
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a, double b) { return a - Dual<NP>(b); }
template <int NP> inline Dual<NP> operator*(const Dual<NP>& a, double b) { return a * Dual<NP>(b); }
template <int NP> inline Dual<NP> operator/(const Dual<NP>& a, double b) { return a / Dual<NP>(b); }
template <int NP> inline Dual<NP>& operator+=(Dual<NP>& a, const Dual<NP>& b) { a = a + b; return a; }
template <int NP> inline Dual<NP>& operator-=(Dual<NP>& a, const Dual<NP>& b) { a = a - b; return a; }
template <int NP> inline Dual<NP>& operator*=(Dual<NP>& a, const Dual<NP>& b) { a = a * b; return a; }
template <int NP> inline Dual<NP>& operator/=(Dual<NP>& a, const Dual<NP>& b) { a = a / b; return a; }
template <int NP> inline bool operator<(const Dual<NP>& a, const Dual<NP>& b) { return a.v < b.v; }
template <int NP> inline bool operator>(const Dual<NP>& a, const Dual<NP>& b) { return a.v > b.v; }
template <int NP> inline bool operator==(const Dual<NP>& a, const Dual<NP>& b) { return a.v == b.v; }
template <int NP> inline double value(const Dual<NP>& x) { return x.v; }
template <int NP> inline Dual<NP> cond_gt(const Dual<NP>& l, const Dual<NP>& r, const Dual<NP>& a, const Dual<NP>& b) { return l.v > r.v ? a : b; }
template <int NP> inline Dual<NP> cond_lt(const Dual<NP>& l, const Dual<NP>& r, const Dual<NP>& a, const Dual<NP>& b) { return l.v < r.v ? a : b; }
template <int NP> inline Dual<NP> cond_eq(const Dual<NP>& l, const Dual<NP>& r, const Dual<NP>& a, const Dual<NP>& b) { return l.v == r.v ? a : b; }
ORC_DUAL_UN(s_abs, std::fabs(x.v), (x.v > 0 ? 1.0 : (x.v < 0 ? -1.0 : 0.0)))
ORC_DUAL_UN(s_sqrt, std::sqrt(x.v), 0.5 / std::sqrt(x.v))
ORC_DUAL_UN(s_sin, std::sin(x.v), std::cos(x.v))
ORC_DUAL_UN(s_cos, std::cos(x.v), -std::sin(x.v))
ORC_DUAL_UN(s_tan, std::tan(x.v), 1.0 + std::tan(x.v) * std::tan(x.v))
ORC_DUAL_UN(s_tanh, std::tanh(x.v), 1.0 - std::tanh(x.v) * std::tanh(x.v))
ORC_DUAL_UN(s_exp, std::exp(x.v), std::exp(x.v))
ORC_DUAL_UN(s_acos, std::acos(x.v), -1.0 / std::sqrt(1.0 - x.v * x.v))
template <int NP> inline Dual<NP> s_atan2(const Dual<NP>& y, const Dual<NP>& x) {
    Dual<NP> r; r.v = std::atan2(y.v, x.v); const double d = x.v * x.v + y.v * y.v;
    for (int i = 0; i < NP; i++) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / d; return r;
}
template <int NP> inline Dual<NP> s_pow(const Dual<NP>& x, const Dual<NP>& y) {
    Dual<NP> r; r.v = std::pow(x.v, y.v);
    const double dx = (x.v == 0.0) ? 0.0 : y.v * r.v / x.v;
    const double dy = (x.v > 0.0) ? r.v * std::log(x.v) : 0.0;
    for (int i = 0; i < NP; i++) r.d[i] = dx * x.d[i] + dy * y.d[i]; return r;
}

// ------------------------------------------------------------------ op counter (for SURVEY §8d work figures)
struct OpCount {
    long long addsub = 0, mul = 0, div = 0, tanh = 0, sqrt = 0, trig = 0, exp = 0, pow = 0, other = 0;
};
inline OpCount*& active_counter() { static thread_local OpCount* c = nullptr; return c; }
struct Cnt {
    double v;
    Cnt() : v(0) {}
    Cnt(double x) : v(x) {}
    Cnt(int x) : v(x) {}
};
inline Cnt operator+(const Cnt& a, const Cnt& b) { active_counter()->addsub++; return Cnt(a.v + b.v); }
inline Cnt operator-(const Cnt& a, const Cnt& b) { active_counter()->addsub++; return Cnt(a.v - b.v); }
inline Cnt operator*(const Cnt& a, const Cnt& b) { active_counter()->mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(const Cnt& a, const Cnt& b) { active_counter()->div++; return Cnt(a.v / b.v); }
inline Cnt operator-(const Cnt& a) { return Cnt(-a.v); }
inline Cnt& operator+=(Cnt& a, const Cnt& b) { a = a + b; return a; }
inline Cnt& operator-=(Cnt& a, const Cnt& b) { a = a - b; return a; }
inline Cnt& operator*=(Cnt& a, const Cnt& b) { a = a * b; return a; }
inline Cnt& operator/=(Cnt& a, const Cnt& b) { a = a / b; return a; }
inline bool operator<(const Cnt& a, const Cnt& b) { return a.v < b.v; }
inline bool operator>(const Cnt& a, const Cnt& b) { return a.v > b.v; }
inline bool operator==(const Cnt& a, const Cnt& b) { return a.v == b.v; }
inline double value(const Cnt& x) { return x.v; }
inline Cnt cond_gt(const Cnt& l, const Cnt& r, const Cnt& a, const Cnt& b) { return l.v > r.v ? a : b; }
inline Cnt cond_lt(const Cnt& l, const Cnt& r, const Cnt& a, const Cnt& b) { return l.v < r.v ? a : b; }
inline Cnt cond_eq(const Cnt& l, const Cnt& r, const Cnt& a, const Cnt& b) { return l.v == r.v ? a : b; }
inline Cnt s_abs(const Cnt& x) { return Cnt(std::fabs(x.v)); }
inline Cnt s_sqrt(const Cnt& x) { active_counter()->sqrt++; return Cnt(std::sqrt(x.v)); }
inline Cnt s_sin(const Cnt& x) { active_counter()->trig++; return Cnt(std::sin(x.v)); }
inline Cnt s_cos(const Cnt& x) { active_counter()->trig++; return Cnt(std::cos(x.v)); }
inline Cnt s_tan(const Cnt& x) { active_counter()->trig++; return Cnt(std::tan(x.v)); }
inline Cnt s_tanh(const Cnt& x) { active_counter()->tanh++; return Cnt(std::tanh(x.v)); }
inline Cnt s_exp(const Cnt& x) { active_counter()->exp++; return Cnt(std::exp(x.v)); }
inline Cnt s_acos(const Cnt& x) { active_counter()->trig++; return Cnt(std::acos(x.v)); }
inline Cnt s_atan2(const Cnt& y, const Cnt& x) { active_counter()->trig++; return Cnt(std::atan2(y.v, x.v)); }
inline Cnt s_pow(const Cnt& x, const Cnt& y) { active_counter()->pow++; return Cnt(std::pow(x.v, y.v)); }

}  // namespace orc
