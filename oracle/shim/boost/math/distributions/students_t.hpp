// oracle/shim/boost/math/distributions/students_t.hpp -- stand-in for the one use in the
// reference (IO::calP, scr/dtpr.cpp:169-174): 2 * cdf(complement(students_t(df), |t|)).
// The P value never reaches the block fit.  Upper tail via the regularised incomplete beta
// function I_x(df/2, 1/2), x = df / (df + t^2), evaluated with a Lentz continued fraction.
#pragma once
#include <cmath>
namespace boost { namespace math {
class students_t {
public:
    explicit students_t(double df) : df_(df) {}
    double degrees_of_freedom() const { return df_; }
private:
    double df_;
};
template <class Dist>
struct complemented2_type { const Dist& dist; double x; complemented2_type(const Dist& d, double v) : dist(d), x(v) {} };
template <class Dist>
inline complemented2_type<Dist> complement(const Dist& d, double x) { return complemented2_type<Dist>(d, x); }
namespace shim_detail {
inline double betacf(double a, double b, double x) {
    const double tiny = 1e-300;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0, c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 500; ++m) {
        const int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < 1e-15) break;
    }
    return h;
}
inline double ibeta(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double bt = std::exp(std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log1p(-x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}
}  // namespace shim_detail
inline double cdf(const complemented2_type<students_t>& c) {
    const double df = c.dist.degrees_of_freedom(), t = c.x;
    const double tail = 0.5 * shim_detail::ibeta(0.5 * df, 0.5, df / (df + t * t));
    return t >= 0 ? tail : 1.0 - tail;
}
}}  // namespace boost::math
