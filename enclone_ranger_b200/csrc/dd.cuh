// dd.cuh — double-double arithmetic for host and device (product code; the oracle has its own
// restatement in oracle/dd.hpp and the two are compared bit-for-bit by tests/).
//
// Operations mirror what the reference performs through the `qd` crate (Cargo.lock:1374-1376)
// in p_at_most_m_distinct_in_sample_of_x_from_n_double (enclone_core/src/join_one.rs:22-43) and
// stirling2_ratio_table_double (enclone_stuff/src/start.rs:50-99): error-free transforms with an
// explicit fma, IEEE-accurate add, product with cross terms, three-quotient division.
// Compile with -fmad=false (device) and -ffp-contract=off (host): contraction would change bits.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define EJB_HD __host__ __device__ __forceinline__
#else
#define EJB_HD inline
#endif

namespace ejb {

struct dd_t {
    double hi, lo;
};

EJB_HD dd_t dd_make(double h, double l = 0.0) {
    dd_t r;
    r.hi = h;
    r.lo = l;
    return r;
}
EJB_HD void dd_two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
EJB_HD void dd_quick_two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    e = b - (s - a);
}
EJB_HD void dd_two_prod(double a, double b, double& p, double& e) {
    p = a * b;
    e = fma(a, b, -p);
}
EJB_HD dd_t dd_add(dd_t a, dd_t b) {
    double s0, e0, s1, e1, s2, e2, r0, r1;
    dd_two_sum(a.hi, b.hi, s0, e0);
    dd_two_sum(a.lo, b.lo, s1, e1);
    dd_quick_two_sum(s0, s1 + e0, s2, e2);
    dd_quick_two_sum(s2, e1 + e2, r0, r1);
    return dd_make(r0, r1);
}
EJB_HD dd_t dd_sub(dd_t a, dd_t b) { return dd_add(a, dd_make(-b.hi, -b.lo)); }
EJB_HD dd_t dd_mul(dd_t a, dd_t b) {
    double p, e, r0, r1;
    dd_two_prod(a.hi, b.hi, p, e);
    e += a.hi * b.lo + a.lo * b.hi;
    dd_quick_two_sum(p, e, r0, r1);
    return dd_make(r0, r1);
}
EJB_HD dd_t dd_mul_d(dd_t a, double b) {
    double p, e, r0, r1;
    dd_two_prod(a.hi, b, p, e);
    e += a.lo * b;
    dd_quick_two_sum(p, e, r0, r1);
    return dd_make(r0, r1);
}
EJB_HD dd_t dd_div(dd_t a, dd_t b) {
    double q1 = a.hi / b.hi;
    dd_t r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    double s, e, t0, t1;
    dd_quick_two_sum(q1, q2, s, e);
    dd_t q = dd_add(dd_make(s, e), dd_make(q3, 0.0));
    dd_quick_two_sum(q.hi, q.lo, t0, t1);
    return dd_make(t0, t1);
}
EJB_HD dd_t dd_powi(dd_t a, int n) {
    if (n == 0) return dd_make(1.0);
    dd_t r = a, s = dd_make(1.0);
    int i = n;
    if (i > 1) {
        while (i > 0) {
            if (i % 2 == 1) s = dd_mul(s, r);
            i /= 2;
            if (i > 0) r = dd_mul(r, r);
        }
    } else {
        s = r;
    }
    return s;
}
EJB_HD bool dd_lt(dd_t a, dd_t b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

}  // namespace ejb
