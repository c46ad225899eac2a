// refshim/cppad/cppad.hpp — TEST INFRASTRUCTURE ONLY (oracle/README.md).
//
// Stand-in for <cppad/cppad.hpp> that lets the reference's OWN first-party sources (read in place from /root/reference, never
// copied) compile in this image, where CppAD is absent.  It implements exactly the part of CppAD's public interface those
// sources use, with CppAD's documented semantics:
//   * AD<double>: operator overloading scalar; while a recording is active (between Independent() and the ADFun constructor
//     on the same thread) every operation whose operands depend on the independents appends one node to the thread's tape
//   * CondExpOp(left, right, if_true, if_false): the comparison is on VALUES; with recording, the selected branch alone is
//     differentiated (CppAD records a conditional; for a first-order Reverse at the recording point this is identical)
//   * abs: derivative sign(x) with sign(0) = 0;  pow(x, y): derivative w.r.t. both arguments;  atan2, sqrt, sin, cos, tan,
//     tanh, exp, acos, log: the usual partials
//   * ADFun<double> f(x, y); f.Reverse(1, w) returns w^T dy/dx at the recorded point (first order only)
//   * thread_alloc::parallel_setup / parallel_ad: accepted, no-ops (the tape is thread_local)
// What it is NOT: CppAD (no forward mode, no higher orders, no tape optimisation, no re-evaluation at other points).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace CppAD {

namespace shim {
struct Node { int32_t a, b; double da, db; };
struct Tape {
    std::vector<Node> nodes;
    bool recording = false;
    std::size_t hint = 0;  // size of the thread's previous recording: reserved up front by the next Independent()
    int push(int a, double da, int b, double db) { nodes.push_back(Node{a, b, da, db}); return (int)nodes.size() - 1; }
};
// one tape per thread, heap-allocated and never destroyed: no thread-exit destructor runs inside a dlopen'ed library
inline Tape& tape() { static thread_local Tape* t = nullptr; if (!t) t = new Tape(); return *t; }
}  // namespace shim

template <class Base> class AD;

template <> class AD<double> {
public:
    double v;
    int id;  // node on the thread's tape, or -1 (a parameter in CppAD's vocabulary)
    AD() : v(0.0), id(-1) {}
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> AD(T x) : v((double)x), id(-1) {}
    AD(double x, int i) : v(x), id(i) {}
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> AD& operator=(T x) { v = (double)x; id = -1; return *this; }
    AD& operator+=(const AD& o);
    AD& operator-=(const AD& o);
    AD& operator*=(const AD& o);
    AD& operator/=(const AD& o);
    AD operator-() const;
    AD operator+() const { return *this; }
};
typedef AD<double> ADd;

namespace shim {
inline ADd mk1(double v, const ADd& a, double da) {
    if (a.id < 0 || !tape().recording) return ADd(v);
    return ADd(v, tape().push(a.id, da, -1, 0.0));
}
inline ADd mk2(double v, const ADd& a, double da, const ADd& b, double db) {
    if ((a.id < 0 && b.id < 0) || !tape().recording) return ADd(v);
    return ADd(v, tape().push(a.id, da, b.id, db));
}
}  // namespace shim

inline ADd operator+(const ADd& a, const ADd& b) { return shim::mk2(a.v + b.v, a, 1.0, b, 1.0); }
inline ADd operator-(const ADd& a, const ADd& b) { return shim::mk2(a.v - b.v, a, 1.0, b, -1.0); }
inline ADd operator*(const ADd& a, const ADd& b) { return shim::mk2(a.v * b.v, a, b.v, b, a.v); }
inline ADd operator/(const ADd& a, const ADd& b) { const double q = a.v / b.v; return shim::mk2(q, a, 1.0 / b.v, b, -q / b.v); }
inline ADd ADd::operator-() const { return shim::mk1(-v, *this, -1.0); }
inline ADd& ADd::operator+=(const ADd& o) { *this = *this + o; return *this; }
inline ADd& ADd::operator-=(const ADd& o) { *this = *this - o; return *this; }
inline ADd& ADd::operator*=(const ADd& o) { *this = *this * o; return *this; }
inline ADd& ADd::operator/=(const ADd& o) { *this = *this / o; return *this; }

#define REFSHIM_MIXED(op)                                                                                                         \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline ADd operator op(const ADd& a, T b) { return a op ADd(b); } \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline ADd operator op(T a, const ADd& b) { return ADd(a) op b; }
REFSHIM_MIXED(+)
REFSHIM_MIXED(-)
REFSHIM_MIXED(*)
REFSHIM_MIXED(/)
#undef REFSHIM_MIXED
#define REFSHIM_CMP(op)                                                                                                           \
    inline bool operator op(const ADd& a, const ADd& b) { return a.v op b.v; }                                                    \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline bool operator op(const ADd& a, T b) { return a.v op (double)b; } \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline bool operator op(T a, const ADd& b) { return (double)a op b.v; }
REFSHIM_CMP(<)
REFSHIM_CMP(>)
REFSHIM_CMP(<=)
REFSHIM_CMP(>=)
REFSHIM_CMP(==)
REFSHIM_CMP(!=)
#undef REFSHIM_CMP

inline std::ostream& operator<<(std::ostream& os, const ADd& x) { return os << x.v; }
inline std::istream& operator>>(std::istream& is, ADd& x) { double v = 0; is >> v; x = ADd(v); return is; }

inline double Value(const ADd& x) { return x.v; }
inline double Value(double x) { return x; }

// ---- elementary functions
inline ADd abs(const ADd& x) { return shim::mk1(std::fabs(x.v), x, x.v > 0 ? 1.0 : (x.v < 0 ? -1.0 : 0.0)); }
inline ADd fabs(const ADd& x) { return abs(x); }
inline ADd sqrt(const ADd& x) { const double r = std::sqrt(x.v); return shim::mk1(r, x, 0.5 / r); }
inline ADd sin(const ADd& x) { return shim::mk1(std::sin(x.v), x, std::cos(x.v)); }
inline ADd cos(const ADd& x) { return shim::mk1(std::cos(x.v), x, -std::sin(x.v)); }
inline ADd tan(const ADd& x) { const double t = std::tan(x.v); return shim::mk1(t, x, 1.0 + t * t); }
inline ADd tanh(const ADd& x) { const double t = std::tanh(x.v); return shim::mk1(t, x, 1.0 - t * t); }
inline ADd exp(const ADd& x) { const double e = std::exp(x.v); return shim::mk1(e, x, e); }
inline ADd log(const ADd& x) { return shim::mk1(std::log(x.v), x, 1.0 / x.v); }
inline ADd acos(const ADd& x) { return shim::mk1(std::acos(x.v), x, -1.0 / std::sqrt(1.0 - x.v * x.v)); }
inline ADd asin(const ADd& x) { return shim::mk1(std::asin(x.v), x, 1.0 / std::sqrt(1.0 - x.v * x.v)); }
inline ADd atan(const ADd& x) { return shim::mk1(std::atan(x.v), x, 1.0 / (1.0 + x.v * x.v)); }
inline ADd atan2(const ADd& y, const ADd& x) {
    const double d = x.v * x.v + y.v * y.v;
    return shim::mk2(std::atan2(y.v, x.v), y, x.v / d, x, -y.v / d);
}
inline ADd pow(const ADd& x, const ADd& y) {
    const double p = std::pow(x.v, y.v);
    const double dx = (x.v == 0.0) ? 0.0 : y.v * p / x.v;
    const double dy = (x.v > 0.0) ? p * std::log(x.v) : 0.0;
    return shim::mk2(p, x, dx, y, dy);
}
template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline ADd pow(const ADd& x, T y) { return pow(x, ADd(y)); }
template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline ADd pow(T x, const ADd& y) { return pow(ADd(x), y); }
// plain floating point through the CppAD namespace (the reference instantiates some templates on double as well)
inline double abs(double x) { return std::fabs(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double tan(double x) { return std::tan(x); }
inline double tanh(double x) { return std::tanh(x); }
inline double exp(double x) { return std::exp(x); }
inline double acos(double x) { return std::acos(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double pow(double x, double y) { return std::pow(x, y); }

// ---- conditional expressions (comparison on values; the selected branch carries the dependence)
#define REFSHIM_CONDEXP(name, op)                                                                                  \
    inline ADd name(const ADd& l, const ADd& r, const ADd& t, const ADd& f) { return (l.v op r.v) ? t : f; }        \
    inline double name(double l, double r, double t, double f) { return (l op r) ? t : f; }
REFSHIM_CONDEXP(CondExpLt, <)
REFSHIM_CONDEXP(CondExpLe, <=)
REFSHIM_CONDEXP(CondExpEq, ==)
REFSHIM_CONDEXP(CondExpGe, >=)
REFSHIM_CONDEXP(CondExpGt, >)
#undef REFSHIM_CONDEXP

// ---- recording
template <class VectorAD> void Independent(VectorAD& x) {
    shim::Tape& t = shim::tape();
    t.nodes.clear();
    if (t.hint) t.nodes.reserve(t.hint + t.hint / 8);
    t.recording = true;
    for (std::size_t i = 0; i < (std::size_t)x.size(); i++) x[i] = ADd(x[i].v, t.push(-1, 0.0, -1, 0.0));
}

template <class Base> class ADFun;
template <> class ADFun<double> {
public:
    ADFun() {}
    // stops the recording and takes the tape over
    template <class VectorAD> ADFun(const VectorAD& x, const VectorAD& y) {
        shim::Tape& t = shim::tape();
        t.hint = t.nodes.size();
        nodes_.swap(t.nodes);
        t.nodes.clear();
        t.recording = false;
        for (std::size_t i = 0; i < (std::size_t)x.size(); i++) xid_.push_back(x[i].id);
        for (std::size_t i = 0; i < (std::size_t)y.size(); i++) yid_.push_back(y[i].id);
    }
    std::size_t Domain() const { return xid_.size(); }
    std::size_t Range() const { return yid_.size(); }
    std::size_t size_var() const { return nodes_.size(); }
    // first-order reverse sweep: returns w^T dy/dx
    template <class VectorD> VectorD Reverse(std::size_t order, const VectorD& w) {
        if (order != 1) throw std::runtime_error("refshim ADFun::Reverse: first order only");
        std::vector<double> adj(nodes_.size(), 0.0);
        for (std::size_t i = 0; i < yid_.size(); i++) if (yid_[i] >= 0) adj[yid_[i]] += w[i];
        for (int i = (int)nodes_.size() - 1; i >= 0; --i) {
            const double a = adj[i];
            if (a == 0.0) continue;
            const shim::Node& n = nodes_[i];
            if (n.a >= 0) adj[n.a] += a * n.da;
            if (n.b >= 0) adj[n.b] += a * n.db;
        }
        VectorD g(xid_.size());
        for (std::size_t i = 0; i < xid_.size(); i++) g[i] = xid_[i] >= 0 ? adj[xid_[i]] : 0.0;
        return g;
    }
private:
    std::vector<shim::Node> nodes_;
    std::vector<int> xid_, yid_;
};

// ---- multi-threading set-up calls of the reference's Trainer (the shim's tape is thread_local: nothing to do)
struct thread_alloc {
    static void parallel_setup(std::size_t, bool (*)(void), std::size_t (*)(void)) {}
    static void hold_memory(bool) {}
};
template <class T> void parallel_ad() {}

inline std::string to_string(const ADd& x) { return std::to_string(x.v); }

}  // namespace CppAD

#include <stdexcept>
