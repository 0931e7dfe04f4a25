// Jorion's Bayes-Stein combination (portfolio_calculations.py:851-895) from y = C^-1 t and z = C^-1 1, C = (m-1) V_hat
// the centred Gram of the m excess returns.  Shared by the two-right-hand-side Cholesky kernel (base windows) and the
// chain kernel (windows solved relative to a base).
//   V_bar = kappa C, kappa = m / ((m-N-2)(m-1))  (:876-879);  mu_g = 1'V_bar^-1 mu_hat / 1'V_bar^-1 1 (:882),
//   q = (mu_hat - mu_g 1)' V_bar^-1 (mu_hat - mu_g 1),  lambda = (N+2)/q (:885),  v = (N+2)/((N+2) + m q) (:887),
//   V_PJ = a V_bar + b 11', a = 1 + 1/(m+lambda), b = lambda / (m (m+1+lambda) 1'V_bar^-1 1) (:888),
//   mu_PJ = (1-v) mu_hat + v mu_g 1 (:889),  nu = V_PJ^-1 mu_PJ (:891-893) by Sherman-Morrison = c_y y + c_z z.
#pragma once

namespace bp {

struct JorionCoef {
    double mu_g, lambda, v, q, one_vinv_one;
    double c_y, c_z;          // nu_j = c_y * y_j + c_z * z_j
};

// sy = 1'y, sz = 1'z, ty = t'y
__device__ __forceinline__ JorionCoef jorion_coefficients(double sy, double sz, double ty, double m, double Nd) {
    JorionCoef c;
    const double kappa = m / ((m - Nd - 2.0) * (m - 1.0));
    const double sym = sy / m;                               // 1'C^-1 mu_hat
    c.mu_g = sym / sz;
    c.q = (ty / (m * m) - sym * sym / sz) / kappa;
    c.lambda = (Nd + 2.0) / c.q;
    c.v = (Nd + 2.0) / ((Nd + 2.0) + m * c.q);
    const double a = 1.0 + 1.0 / (m + c.lambda);
    c.one_vinv_one = sz / kappa;
    const double b = c.lambda / (m * (m + 1.0 + c.lambda)) / c.one_vinv_one;
    const double iak = 1.0 / (a * kappa);
    const double one_r = sym * iak;                          // 1'(aV_bar)^-1 mu_PJ  (since mu_g 1'z = 1'y/m)
    const double one_s = sz * iak;                           // 1'(aV_bar)^-1 1
    const double corr = b * one_r / (1.0 + b * one_s);
    c.c_y = (1.0 - c.v) * iak / m;
    c.c_z = (c.v * c.mu_g - corr) * iak;
    return c;
}

}  // namespace bp
