// Host-side 3x3 algebra of the ExpAns kernel in a DEFINED operation order (compile with -ffp-contract=off).
// Mirrors oracle/gpss_oracle.py: rot_matrix, sig_inv, s_matrices.  Citations are into /root/reference.
#pragma once
#include <cmath>

namespace gpss {

struct Rot3 { double r[3][3]; };

// Rot (Kernel.cpp:1402-1410); alpha=AngleX, beta=AngleY, teta=AngleZ (Kernel.cpp:866-868)
inline Rot3 rot_matrix(double alpha, double beta, double teta)
{
  const double sa = std::sin(alpha), ca = std::cos(alpha);
  const double sb = std::sin(beta), cb = std::cos(beta);
  const double st = std::sin(teta), ct = std::cos(teta);
  Rot3 R;
  R.r[0][0] = ca * ct + sa * sb * st;
  R.r[0][1] = -sa * ct + ca * sb * st;
  R.r[0][2] = -cb * st;
  R.r[1][0] = sa * cb;
  R.r[1][1] = ca * cb;
  R.r[1][2] = sb;
  R.r[2][0] = ca * st - sa * sb * ct;
  R.r[2][1] = -sa * st - ca * sb * ct;
  R.r[2][2] = cb * ct;
  return R;
}

// sigInv = Rot*lambda*Rot.t() (Kernel.cpp:1417-1425): S(i,j) = (t0*R(j,0) + t1*R(j,1)) + t2*R(j,2), t_k = R(i,k)*l_k
inline void sig_inv(const double theta[10], double S[9])
{
  const Rot3 R = rot_matrix(theta[0], theta[2], theta[4]);
  const double lam[3] = {theta[1], theta[3], theta[5]};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const double t0 = R.r[i][0] * lam[0];
      const double t1 = R.r[i][1] * lam[1];
      const double t2 = R.r[i][2] * lam[2];
      S[i * 3 + j] = (t0 * R.r[j][0] + t1 * R.r[j][1]) + t2 * R.r[j][2];
    }
}

// MahaDist centre (Kernel.cpp:1391-1392) from column sums accumulated in row order
inline void maha_centre(long n, const double* sums1, long m, const double* sums2, double* c, int d = 3)
{
  for (int j = 0; j < d; j++) {
    const double mX1 = ((double)n / (double)(n + m)) * sums1[j] / (double)n;
    c[j] = ((double)m / (double)(n + m)) * sums2[j] / (double)m + mX1;
  }
}

inline void seq_colsums(const double* X_colmajor, long n, double* sums, int d = 3)
{
  for (int j = 0; j < d; j++) {
    double acc = 0.0;
    const double* col = X_colmajor + (long)j * n;
    for (long i = 0; i < n; i++) acc += col[i];
    sums[j] = acc;
  }
}

// M_p = S % S_p for p = AngleX, iWx, AngleY, iWy, AngleZ, iWz with S and S_p written out entry by entry
// as Kern_ExpAnisotropic::getGradients does (Kernel.cpp:955-1166), INCLUDING the (0,0) quirk: the z term
// of S_angle(0,0) carries no factor 2 (Kernel.cpp:1003-1011).
inline void grad_M_matrices(const double theta[10], double M[6][9])
{
  const double al = theta[0], be = theta[2], te = theta[4];
  const double l[3] = {theta[1], theta[3], theta[5]};
  const double sa = std::sin(al), ca = std::cos(al), sb = std::sin(be), cb = std::cos(be), st = std::sin(te), ct = std::cos(te);
  const Rot3 Rm = rot_matrix(al, be, te);
  const double (*R)[3] = Rm.r;
  double D[3][3][3];   // D[0]=dRot/dalpha, D[1]=dRot/dbeta, D[2]=dRot/dteta (Kernel.cpp:955-998)
  D[0][0][0] = -sa * ct + ca * sb * st;  D[1][0][0] = sa * cb * st;   D[2][0][0] = -ca * st + sa * sb * ct;
  D[0][0][1] = -ca * ct - sa * sb * st;  D[1][0][1] = ca * cb * st;   D[2][0][1] = sa * st + ca * sb * ct;
  D[0][0][2] = 0.0;                      D[1][0][2] = sb * st;        D[2][0][2] = -cb * ct;
  D[0][1][0] = ca * cb;                  D[1][1][0] = -sa * sb;       D[2][1][0] = 0.0;
  D[0][1][1] = -sa * cb;                 D[1][1][1] = -ca * sb;       D[2][1][1] = 0.0;
  D[0][1][2] = 0.0;                      D[1][1][2] = cb;             D[2][1][2] = 0.0;
  D[0][2][0] = -sa * st - ca * sb * ct;  D[1][2][0] = -sa * cb * ct;  D[2][2][0] = ca * ct + sa * sb * st;
  D[0][2][1] = -ca * st + sa * sb * ct;  D[1][2][1] = -ca * cb * ct;  D[2][2][1] = -sa * ct + ca * sb * st;
  D[0][2][2] = 0.0;                      D[1][2][2] = -sb * ct;       D[2][2][2] = -cb * st;

  double S[3][3], Sd[3][3][3], SL[3][3][3];
  S[0][0] = l[0] * (R[0][0] * R[0][0]) + l[1] * (R[0][1] * R[0][1]) + l[2] * (R[0][2] * R[0][2]);
  for (int a = 0; a < 3; a++)
    Sd[a][0][0] = l[0] * 2 * R[0][0] * D[a][0][0] + l[1] * 2 * R[0][1] * D[a][0][1] + l[2] * R[0][2] * D[a][0][2];
  for (int q = 0; q < 3; q++) SL[q][0][0] = R[0][q] * R[0][q];
  static const int IJ[5][2] = {{0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
  for (int e = 0; e < 5; e++) {
    const int i = IJ[e][0], j = IJ[e][1];
    S[i][j] = l[0] * R[i][0] * R[j][0] + l[1] * R[i][1] * R[j][1] + l[2] * R[i][2] * R[j][2];
    for (int a = 0; a < 3; a++)
      Sd[a][i][j] = l[0] * D[a][i][0] * R[j][0] + l[0] * R[i][0] * D[a][j][0]
                  + l[1] * D[a][i][1] * R[j][1] + l[1] * R[i][1] * D[a][j][1]
                  + l[2] * D[a][i][2] * R[j][2] + l[2] * R[i][2] * D[a][j][2];
    for (int q = 0; q < 3; q++) SL[q][i][j] = R[i][q] * R[j][q];
  }
  static const int LO[3][2] = {{1, 0}, {2, 0}, {2, 1}};
  for (int e = 0; e < 3; e++) {
    const int i = LO[e][0], j = LO[e][1];
    S[i][j] = S[j][i];
    for (int a = 0; a < 3; a++) Sd[a][i][j] = Sd[a][j][i];
    for (int q = 0; q < 3; q++) SL[q][i][j] = SL[q][j][i];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      M[0][i * 3 + j] = S[i][j] * Sd[0][i][j];
      M[1][i * 3 + j] = S[i][j] * SL[0][i][j];
      M[2][i * 3 + j] = S[i][j] * Sd[1][i][j];
      M[3][i * 3 + j] = S[i][j] * SL[1][i][j];
      M[4][i * 3 + j] = S[i][j] * Sd[2][i][j];
      M[5][i * 3 + j] = S[i][j] * SL[2][i][j];
    }
}

// Combine the device reductions into the reference's g[0..9].
//   red = {T00,T01,T02,T11,T12,T22, V0,V1,V2, G6, TR, QK, RK},  s3 = sum((y-f)^2/sn2 - 1)
inline void combine_gradient(const double theta[10], const double red[13], double s3, double g[10], int dim = 3, long n = 1)
{
  double M[6][9];
  grad_M_matrices(theta, M);
  const double T[3][3] = {{red[0], red[1], red[2]}, {red[1], red[3], red[4]}, {red[2], red[4], red[5]}};
  for (int p = 0; p < 6; p++) {
    double qv = 0.0, mt = 0.0;
    for (int k = 0; k < 3; k++) {
      const double rho = (M[p][k * 3 + 0] + M[p][k * 3 + 1]) + M[p][k * 3 + 2];
      qv += rho * red[6 + k];
      for (int q = 0; q < 3; q++) mt += M[p][k * 3 + q] * T[k][q];
    }
    g[p] = 4.0 * qv - 4.0 * mt;                 // Kernel.cpp:1192-1233 in reduced form
  }
  g[6] = 2.0 * red[9] * theta[6];               // Kernel.cpp:1239-1242
  // 3 columns: 0 (Kernel.cpp:1256-1257).  4 columns (Kernel.cpp:1246-1255): dhp = -2 RColon' * Di2(:) / n with
  // Di2_ij = 2 (x_i3 - x_j3)^2 and [quirk] RColon still holding exp(-s) from the Sigma gradient (:1240) -- QW does not
  // enter; RK is the sum over i > j, so the ordered-pair sum is 2 RK.
  g[7] = (dim == 4) ? (-2.0 * (2.0 * (2.0 * red[12]))) / (double)n : 0.0;
  g[8] = red[10];                               // Kern_Bias::getGradients = trace(QW) (Kernel.cpp:370-377)
  g[9] = -1.0 * (0.5 * red[11]) * (2.0 / theta[9]) - s3;   // GP_Utils.cpp:1226 with dW = 0.5*rowsum(Q%K) (:1206)
}

// ---------------------------------------------------------------------------------------------------
// The isotropic members of the kernel family: Hyb{Exp, Bias} and Hyb{RBF, Bias}.  theta layouts in the C ABI's 10-array:
//   kind 0 ExpAns  {AngleX, iWx, AngleY, iWy, AngleZ, iWz, Sigma, iWR, Sigma_Bias, sn2}          10 parameters
//   kind 1 Exp     {Hayper_Euc_Exp, Sigma_Exp, Sigma_Bias, sn2}                                  4
//   kind 2 RBF     {Hayper_Euc_RBF, inverseWidth_RBF, Sigma_RBF, Sigma_Bias, sn2}                5
// ---------------------------------------------------------------------------------------------------
inline int kernel_npar(int kind) { return kind == 0 ? 10 : kind == 1 ? 4 : 5; }
inline double theta_sigma(int kind, const double* theta) { return kind == 0 ? theta[6] : theta[kernel_npar(kind) - 3]; }
inline double theta_bias(int kind, const double* theta) { return theta[kernel_npar(kind) - 2]; }
inline double theta_sn2(int kind, const double* theta) { return theta[kernel_npar(kind) - 1]; }

// red as combine_gradient (slots 9 = A, 10 = tr QW, 11 = sum Q o K, 12 = B of grad_pass_kernel's isotropic branch)
inline void combine_gradient_iso(int kind, const double* theta, const double red[13], double s3, double* g)
{
  const double sig = theta_sigma(kind, theta), var2 = sig * sig;
  const int np = kernel_npar(kind);
  if (kind == 1) {
    g[0] = var2 * red[12];                      // sum((var2 QW) % dk % D2), Kernel.cpp:668-681
    g[1] = red[9] * sig;                        // sum(KD2 % (QW % KD2)) * Sigma_Exp, Kernel.cpp:683-691
  } else {
    const double w = theta[1];
    g[0] = (-2.0 * ((var2 * (-w / 2)) * red[9])) / 2;      // Kernel.cpp:517-525, 537
    g[1] = ((-0.5 * var2) * red[9]) / 2;                   // Kernel.cpp:527-528, 538
    g[2] = (((red[12] * sig) + (red[12] * sig)) * sig) / 2;   // Kernel.cpp:530-536, 539
  }
  g[np - 2] = red[10];                                         // Kern_Bias: trace(QW)
  g[np - 1] = -1.0 * (0.5 * red[11]) * (2.0 / theta[np - 1]) - s3;
}

}  // namespace gpss
