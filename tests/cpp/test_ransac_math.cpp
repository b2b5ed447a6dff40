// Host-side unit test of gtsfm_b200/csrc/ransac_math.cuh (the same source the CUDA verifier compiles for the device).
// Build: g++ -O2 -std=c++17 -x c++ tests/cpp/test_ransac_math.cpp -o /tmp/test_ransac_math && /tmp/test_ransac_math
#include <stdio.h>
#include <stdlib.h>

#include <random>

#include "../../gtsfm_b200/csrc/ransac_math.cuh"

using namespace rmath;

static void rot(double ax, double ay, double az, double* R) {
  double cx = cos(ax), sx = sin(ax), cy = cos(ay), sy = sin(ay), cz = cos(az), sz = sin(az);
  double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  double T[9];
  mat3_mul(Rz, Ry, T);
  mat3_mul(T, Rx, R);
}

int main() {
  std::mt19937_64 gen(7);
  std::uniform_real_distribution<double> U(-1, 1);
  int fails = 0, trials = 300, found = 0, nsol_total = 0;
  double worst = 0;
  for (int tr = 0; tr < trials; ++tr) {
    double R[9], t[3] = {U(gen), U(gen), 0.3 * U(gen)};
    rot(0.3 * U(gen), 0.3 * U(gen), 0.3 * U(gen), R);
    double tn = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    for (int i = 0; i < 3; ++i) t[i] /= tn;
    double tx[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0}, E[9];
    mat3_mul(tx, R, E);
    double x1[5][2], x2[5][2];
    for (int p = 0; p < 5; ++p) {
      double X[3] = {U(gen), 0.8 * U(gen), 4 + 3 * (U(gen) + 1)};
      x1[p][0] = X[0] / X[2], x1[p][1] = X[1] / X[2];
      double Y[3];
      for (int i = 0; i < 3; ++i) Y[i] = R[i * 3] * X[0] + R[i * 3 + 1] * X[1] + R[i * 3 + 2] * X[2] + t[i];
      x2[p][0] = Y[0] / Y[2], x2[p][1] = Y[1] / Y[2];
    }
    double sols[10][9];
    int n = fivept_solve(x1, x2, sols);
    nsol_total += n;
    double en = 0;
    for (int i = 0; i < 9; ++i) en += E[i] * E[i];
    en = sqrt(en);
    double best = 1e9;
    for (int s = 0; s < n; ++s) {
      double dp = 0, dm = 0;
      for (int i = 0; i < 9; ++i) {
        dp += (sols[s][i] - E[i] / en) * (sols[s][i] - E[i] / en);
        dm += (sols[s][i] + E[i] / en) * (sols[s][i] + E[i] / en);
      }
      double d = sqrt(dp < dm ? dp : dm);
      if (d < best) best = d;
      // every returned solution must satisfy the 5 epipolar constraints
      for (int p = 0; p < 5; ++p)
        if (sampson_sq(sols[s], x1[p][0], x1[p][1], x2[p][0], x2[p][1]) > 1e-12) { ++fails; break; }
    }
    if (best < 1e-6) ++found;
    if (best < 1e8 && best > worst && best < 1e-6) worst = best;
    // pose recovery from the true E
    double R1[9], R2[9], tt[3];
    decompose_E(E, R1, R2, tt);
    const double* Rs[4] = {R1, R1, R2, R2};
    int votes[4] = {0, 0, 0, 0};
    for (int c = 0; c < 4; ++c) {
      double tc[3] = {(c & 1) ? -tt[0] : tt[0], (c & 1) ? -tt[1] : tt[1], (c & 1) ? -tt[2] : tt[2]};
      for (int p = 0; p < 5; ++p) votes[c] += cheirality_ok(Rs[c], tc, x1[p][0], x1[p][1], x2[p][0], x2[p][1], 50.0);
    }
    int bc = 0;
    for (int c = 1; c < 4; ++c) if (votes[c] > votes[bc]) bc = c;
    double dR = 0;
    for (int i = 0; i < 9; ++i) dR += fabs(Rs[bc][i] - R[i]);
    double sgn = (bc & 1) ? -1 : 1;
    double dt = fabs(sgn * tt[0] - t[0]) + fabs(sgn * tt[1] - t[1]) + fabs(sgn * tt[2] - t[2]);
    if (votes[bc] != 5 || dR > 1e-6 || dt > 1e-6) { ++fails; printf("pose fail trial %d votes %d dR %g dt %g\n", tr, votes[bc], dR, dt); }
  }
  printf("fivept: true E recovered in %d/%d trials, avg %.2f solutions, worst err %.2e, constraint/pose fails %d\n", found, trials,
         (double)nsol_total / trials, worst, fails);
  // svd3 check
  double M[9] = {1, 2, 3, 4, 5, 6, 7, 8, 10}, Um[9], s[3], V[9], rec[9];
  svd3(M, Um, s, V);
  double err = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      rec[i * 3 + j] = 0;
      for (int k = 0; k < 3; ++k) rec[i * 3 + j] += Um[i * 3 + k] * s[k] * V[j * 3 + k];
      err += fabs(rec[i * 3 + j] - M[i * 3 + j]);
    }
  printf("svd3 reconstruction err %.2e\n", err);
  return (found >= trials * 0.97 && fails == 0 && err < 1e-9) ? 0 : 1;
}
