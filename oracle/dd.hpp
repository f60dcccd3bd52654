// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by
// the product library (enclone_ranger_b200/csrc); only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
//
// dd.hpp — double-double arithmetic restating the operations the reference performs through
// the third-party crate `qd` 0.2.0-alpha (git https://github.com/Barandis/qd, rev
// 0fb276d70346f11f4b2a5b30568d8a26d0dd88df; reference Cargo.lock:1374-1376).  The crate source
// is NOT under /root/reference, so this follows the published QD algorithms (Hida, Li, Bailey,
// "Library for double-double and quad-double arithmetic", 2007) that `qd` ports:
// error-free transforms two_sum / quick_two_sum / two_prod(fma), IEEE-accurate add,
// product with cross terms, three-quotient division, binary powi.
//
// PARITY UNPINNED at double-double bit level: the reference holds no test of `qd` values on
// this path.  Call sites restated: join_one.rs:28-42 (dd!, *, /, *=, -=, <, f64::from),
// start.rs:52-97 (Double::from(u32), /, *, *=, +, powi).  Every dd! operand there is an
// integer-valued f64, so construction is exact.  The f64 results are pinned by the KATs of
// stirling_numbers/src/lib.rs:261-315 and by an mpmath cross-check (tests/test_oracle_dd.py).
//
// Must be compiled with -ffp-contract=off (Rust never contracts a*b+c).
#pragma once
#include <cmath>

namespace oracle {

struct DD {
    double hi, lo;
    DD() : hi(0.0), lo(0.0) {}
    explicit DD(double x) : hi(x), lo(0.0) {}
    DD(double h, double l) : hi(h), lo(l) {}
};

static inline void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
static inline void quick_two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    e = b - (s - a);
}
static inline void two_prod(double a, double b, double& p, double& e) {
    p = a * b;
    e = std::fma(a, b, -p);
}

static inline DD dd_add(const DD& a, const DD& b) {
    double s0, e0, s1, e1, s2, e2, r0, r1;
    two_sum(a.hi, b.hi, s0, e0);
    two_sum(a.lo, b.lo, s1, e1);
    quick_two_sum(s0, s1 + e0, s2, e2);
    quick_two_sum(s2, e1 + e2, r0, r1);
    return DD(r0, r1);
}
static inline DD dd_neg(const DD& a) { return DD(-a.hi, -a.lo); }
static inline DD dd_sub(const DD& a, const DD& b) { return dd_add(a, dd_neg(b)); }

static inline DD dd_mul(const DD& a, const DD& b) {
    double p, e, r0, r1;
    two_prod(a.hi, b.hi, p, e);
    e += a.hi * b.lo + a.lo * b.hi;
    quick_two_sum(p, e, r0, r1);
    return DD(r0, r1);
}
static inline DD dd_mul_d(const DD& a, double b) {
    double p, e, r0, r1;
    two_prod(a.hi, b, p, e);
    e += a.lo * b;
    quick_two_sum(p, e, r0, r1);
    return DD(r0, r1);
}
static inline DD dd_div(const DD& a, const DD& b) {
    double q1 = a.hi / b.hi;
    DD r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    // renormalise (q1, q2, q3) to two components
    double s, e, t0, t1;
    quick_two_sum(q1, q2, s, e);
    DD q = dd_add(DD(s, e), DD(q3, 0.0));
    quick_two_sum(q.hi, q.lo, t0, t1);
    return DD(t0, t1);
}
static inline DD dd_sqr(const DD& a) { return dd_mul(a, a); }
static inline DD dd_powi(const DD& a, int n) {
    // only non-negative exponents occur (start.rs:90)
    if (n == 0) return DD(1.0);
    DD r = a, s(1.0);
    int i = n;
    if (i > 1) {
        while (i > 0) {
            if (i % 2 == 1) s = dd_mul(s, r);
            i /= 2;
            if (i > 0) r = dd_sqr(r);
        }
    } else {
        s = r;
    }
    return s;
}
static inline bool dd_lt(const DD& a, const DD& b) {
    return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}

}  // namespace oracle
