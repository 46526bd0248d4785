/*
 * oracle/mhd3d_oracle.c  --  TEST INFRASTRUCTURE ONLY (see mhd3d_oracle.h for the rules).
 *
 * CPU restatement, in plain C, of the reference's 3-D MHD step, implementationVersion 0.
 * PARITY PIN: checked bit-for-bit against the unmodified reference binary (oracle/_ref/ppkMHD)
 * through the golden .vti fixtures in tests/golden/ (tests/test_oracle_vs_golden.py).
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared -fopenmp mhd3d_oracle.c -o liboracle.so -lm
 * (-ffp-contract=off: the reference's x86-64 build executes no FMA; keep it that way.)
 */
#include "mhd3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ID = ORC_ID, IP = ORC_IP, IU = ORC_IU, IV = ORC_IV, IW = ORC_IW, IA = ORC_IA, IB = ORC_IB, IC = ORC_IC, NV = ORC_NVAR };
/* src/shared/enums.h: EdgeIndex, EdgeIndex2, EmfIndex */
enum { IRT = 0, IRB = 1, ILT = 2, ILB = 3 };
enum { ILL = 0, IRL = 1, ILR = 2, IRR = 3 };
enum { I_EMFZ = 0, I_EMFY = 1, I_EMFX = 2 };

typedef double state_t[NV];

#define AT(p, i, j, k, v) \
  ((size_t)(i) + (size_t)(p)->isize * ((size_t)(j) + (size_t)(p)->jsize * ((size_t)(k) + (size_t)(p)->ksize * (size_t)(v))))

/* ------------------------------------------------------------------------------------------ */
double orc_parse_float(const char *text, double default_value)
{
  /* src/utils/config/ConfigMap.cpp:37-46 : value goes through strtof, i.e. float precision;
   * the default is itself a float argument. */
  char *end;
  float v = strtof(text, &end);
  return end > text ? (double)v : (double)(float)default_value;
}

void orc_params_finalize(orc_params *p)
{
  /* src/shared/HydroParams.cpp:421-441 and :400-402 */
  if (p->mx < 1) p->mx = 1;
  if (p->my < 1) p->my = 1;
  if (p->mz < 1) p->mz = 1;
  p->isize = p->nx + 2 * p->gw;
  p->jsize = p->ny + 2 * p->gw;
  p->ksize = p->nz + 2 * p->gw;
  p->dx = (p->xmax - p->xmin) / (p->nx * p->mx);
  p->dy = (p->ymax - p->ymin) / (p->ny * p->my);
  p->dz = (p->zmax - p->zmin) / (p->nz * p->mz);
  p->smallp = p->smallc * p->smallc / p->gamma0;
}

long orc_ncells(const orc_params *p) { return (long)p->isize * p->jsize * p->ksize; }

static void load_state(const orc_params *p, const double *A, int i, int j, int k, state_t q)
{
  for (int v = 0; v < NV; ++v) q[v] = A[AT(p, i, j, k, v)];
}
static void store_state(const orc_params *p, double *A, int i, int j, int k, const state_t q)
{
  for (int v = 0; v < NV; ++v) A[AT(p, i, j, k, v)] = q[v];
}

/* ------------------------------------------------------------------------------------------ */
/* Initial conditions                                                                         */
/* ------------------------------------------------------------------------------------------ */
void orc_init_orszag_tang(const orc_params *p, double kt, double *U)
{
  /* src/muscl/MHDInitFunctors3D.h:264-415 : two sweeps over the full array (ghosts included). */
  const double pi = 3.141592653589793238462643383279502884L; /* real_type.h:64 */
  const double twopi = 2 * pi;                               /* real_type.h:75 */
  const int gw = p->gw;
  const double gamma0 = p->gamma0;
  const double B0 = 1.0 / sqrt(4 * pi);
  const double p0 = gamma0 / (4 * pi);
  const double d0 = gamma0 * p0;
  const double v0 = 1.0;

  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) { /* :311-374 */
        double xPos = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double yPos = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double zPos = p->zmin + p->dz / 2 + (k + p->nz * p->pz - gw) * p->dz;
        U[AT(p, i, j, k, ID)] = d0;
        U[AT(p, i, j, k, IU)] = -d0 * v0 * sin(yPos * twopi);
        U[AT(p, i, j, k, IV)] = d0 * v0 * sin(xPos * twopi);
        U[AT(p, i, j, k, IW)] = 0.0;
        U[AT(p, i, j, k, IA)] = -B0 * cos(2 * twopi * kt * (zPos - p->zmin) / (p->zmax - p->zmin)) * sin(yPos * twopi);
        U[AT(p, i, j, k, IB)] = B0 * cos(2 * twopi * kt * (zPos - p->zmin) / (p->zmax - p->zmin)) * sin(2.0 * xPos * twopi);
        U[AT(p, i, j, k, IC)] = 0.0;
        U[AT(p, i, j, k, IP)] = 0.0; /* Kokkos::View is zero-initialised; the energy sweep skips the last i / j */
      }
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) { /* :376-409 */
        if (i < p->isize - 1 && j < p->jsize - 1) {
          double mu = U[AT(p, i, j, k, IU)], mv = U[AT(p, i, j, k, IV)], d = U[AT(p, i, j, k, ID)];
          double sa = U[AT(p, i, j, k, IA)] + U[AT(p, i + 1, j, k, IA)];
          double sb = U[AT(p, i, j, k, IB)] + U[AT(p, i, j + 1, k, IB)];
          U[AT(p, i, j, k, IP)] =
            p0 / (gamma0 - 1.0) + 0.5 * ((mu * mu) / d + (mv * mv) / d + 0.25 * (sa * sa) + 0.25 * (sb * sb));
        }
      }
}

void orc_init_blast(const orc_params *p, double radius, double cx, double cy, double cz, double density_in,
                    double density_out, double pressure_in, double pressure_out, double *U)
{
  /* src/muscl/MHDInitFunctors3D.h:176-252 */
  const int gw = p->gw;
  const double radius2 = radius * radius;
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) {
        double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double z = p->zmin + p->dz / 2 + (k + p->nz * p->pz - gw) * p->dz;
        double d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz);
        int inside = d2 < radius2;
        double a = 0.5, b = 0.5, c = 0.5;
        U[AT(p, i, j, k, ID)] = inside ? density_in : density_out;
        U[AT(p, i, j, k, IU)] = 0.0;
        U[AT(p, i, j, k, IV)] = 0.0;
        U[AT(p, i, j, k, IW)] = 0.0;
        U[AT(p, i, j, k, IA)] = a;
        U[AT(p, i, j, k, IB)] = b;
        U[AT(p, i, j, k, IC)] = c;
        U[AT(p, i, j, k, IP)] = (inside ? pressure_in : pressure_out) / (p->gamma0 - 1.0) + 0.5 * (a * a + b * b + c * c);
      }
}

void orc_init_implode(const orc_params *p, const double outer[8], const double inner[8], int shape, double *U)
{
  /* src/muscl/MHDInitFunctors3D.h:52-143 ; outer/inner = rho, p, u, v, w, Bx, By, Bz (ImplodeParams.h:36-56).
   * The momentum slots receive the VELOCITIES, as in the reference. */
  const int gw = p->gw;
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) {
        double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double z = p->zmin + p->dz / 2 + (k + p->nz * p->pz - gw) * p->dz;
        int tmp;
        if (shape == 1) tmp = x + y + z > 0.5 && x + y + z < 2.5;
        else tmp = x + y + z > (p->xmin + p->xmax) / 2. + p->ymin + p->zmin;
        const double *s = tmp ? outer : inner;
        U[AT(p, i, j, k, ID)] = s[0];
        U[AT(p, i, j, k, IP)] = s[1] / (p->gamma0 - 1.0) + 0.5 * s[0] * (s[2] * s[2] + s[3] * s[3] + s[4] * s[4]) +
                                0.5 * (s[5] * s[5] + s[6] * s[6] + s[7] * s[7]);
        U[AT(p, i, j, k, IU)] = s[2];
        U[AT(p, i, j, k, IV)] = s[3];
        U[AT(p, i, j, k, IW)] = s[4];
        U[AT(p, i, j, k, IA)] = s[5];
        U[AT(p, i, j, k, IB)] = s[6];
        U[AT(p, i, j, k, IC)] = s[7];
      }
}

void orc_init_kelvin_helmholtz(const orc_params *p, double d_in, double d_out, double pressure, double vflow_in,
                               double vflow_out, int mode, double w0, double delta, int sine_robertson, double *U)
{
  /* src/muscl/MHDInitFunctors3D.h:532-614 : the two deterministic perturbations (sine "a la Robertson" :532-573,
   * plain sine :574-612); the random one (:492-530) draws from a per-thread Kokkos pool and is not reproducible. */
  const int gw = p->gw;
  const double z1 = 0.25, z2 = 0.75;
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) {
        double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double z = p->zmin + p->dz / 2 + (k + p->nz * p->pz - gw) * p->dz;
        double d, u, v, w;
        if (sine_robertson) {
          const double rho1 = d_in, rho2 = d_out, v1x = vflow_in, v2x = vflow_out, v1y = vflow_in / 2, v2y = vflow_out / 2;
          const double ramp = 1.0 / (1.0 + exp(2 * (z - z1) / delta)) + 1.0 / (1.0 + exp(2 * (z2 - z) / delta));
          d = rho1 + ramp * (rho2 - rho1);
          u = v1x + ramp * (v2x - v1x);
          v = v1y + ramp * (v2y - v1y);
          w = w0 * sin(mode * M_PI * x) * sin(mode * M_PI * y);
        } else {
          d = (z >= z1 && z <= z2) ? d_in : d_out;
          u = (z >= z1 && z <= z2) ? vflow_in : vflow_out;
          v = 0;
          w = w0 * sin(mode * M_PI * x);
        }
        const double bx = 0.5, by = 0.0, bz = 0.0;
        U[AT(p, i, j, k, ID)] = d;
        U[AT(p, i, j, k, IU)] = d * u;
        U[AT(p, i, j, k, IV)] = d * v;
        U[AT(p, i, j, k, IW)] = d * w;
        U[AT(p, i, j, k, IA)] = bx;
        U[AT(p, i, j, k, IB)] = by;
        U[AT(p, i, j, k, IC)] = bz;
        U[AT(p, i, j, k, IP)] = pressure / (p->gamma0 - 1.0) + 0.5 * d * (u * u + v * v + w * w) + 0.5 * (bx * bx + by * by + bz * bz);
      }
}

void orc_init_rotor(const orc_params *p, double r0, double r1, double u0, double p0, double b0, double *U)
{
  /* src/muscl/MHDInitFunctors3D.h:646-748 (velocities in the momentum slots, as in the reference) */
  const int gw = p->gw;
  const double xCenter = (p->xmax + p->xmin) / 2, yCenter = (p->ymax + p->ymin) / 2;
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) {
        double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double r = sqrt((x - xCenter) * (x - xCenter) + (y - yCenter) * (y - yCenter));
        double f_r = (r1 - r) / (r1 - r0);
        double d, mu, mv;
        if (r <= r0) { d = 10.0; mu = -u0 * (y - yCenter) / r0; mv = u0 * (x - xCenter) / r0; }
        else if (r <= r1) { d = 1 + 9 * f_r; mu = -f_r * u0 * (y - yCenter) / r; mv = f_r * u0 * (x - xCenter) / r; }
        else { d = 1.0; mu = 0.0; mv = 0.0; }
        U[AT(p, i, j, k, ID)] = d;
        U[AT(p, i, j, k, IU)] = mu;
        U[AT(p, i, j, k, IV)] = mv;
        U[AT(p, i, j, k, IW)] = 0.0;
        U[AT(p, i, j, k, IA)] = b0;
        U[AT(p, i, j, k, IB)] = 0.0;
        U[AT(p, i, j, k, IC)] = 0.0;
        U[AT(p, i, j, k, IP)] = p0 / (p->gamma0 - 1.0) + (mu * mu + mv * mv + 0.0 * 0.0) / 2 / d + (b0 * b0) / 2;
      }
}

int orc_init_wave(const orc_params *p, double wave_amplitude, int wave_type, double *U)
{
  /* Linear MHD wave on a rotated axis: src/shared/problems/WaveParams.h:30-150 (eigenvector table, angles, k_par) and
   * src/muscl/MHDInitFunctors3D.h:1034-1308 (vector potential on edges -> face B by a discrete curl -> interior cells).
   * Cells outside the ranges written below stay zero, like the reference's zero-initialised View. Returns -1 for an
   * unknown wave type (the reference exits). */
  double rev[7], wave_V0 = 0.0;
  if (wave_type == 0) {
    const double r[7] = {4.472136e-01, -8.944272e-01, 4.216370e-01, 1.490712e-01, 2.012461e+00, 8.432740e-01, 2.981424e-01};
    memcpy(rev, r, sizeof(r));
  } else if (wave_type == 1) {
    const double r[7] = {0.0, 0.0, -3.333333e-01, 9.428090e-01, 0.0, -3.333333e-01, 9.428090e-01};
    memcpy(rev, r, sizeof(r));
  } else if (wave_type == 2) {
    const double r[7] = {8.944272e-01, -4.472136e-01, -8.432740e-01, -2.981424e-01, 6.708204e-01, -4.216370e-01, -1.490712e-01};
    memcpy(rev, r, sizeof(r));
  } else if (wave_type == 3) {
    const double r[7] = {1.0, 1.0, 0.0, 0.0, 0.5, 0.0, 0.0};
    memcpy(rev, r, sizeof(r));
    wave_V0 = 1.0;
  } else {
    return -1;
  }
  const double Lx = p->xmax - p->xmin, Ly = p->ymax - p->ymin, Lz = p->zmax - p->zmin;
  const double d0 = 1.0, p0 = 1.0 / p->gamma0;
  const double TwoPi = 4.0 * asin(1.0);
  const double ang_3 = atan(Lx / Ly);
  const double sin_a3 = sin(ang_3), cos_a3 = cos(ang_3);
  const double ang_2 = atan(0.5 * (Lx * cos_a3 + Ly * sin_a3) / Lz);
  const double sin_a2 = sin(ang_2), cos_a2 = cos(ang_2);
  const double x1l = Lx * cos_a2 * cos_a3, x2l = Ly * cos_a2 * sin_a3, x3l = Lz * sin_a2;
  double lambda = x1l;
  if (ang_3 != 0.0) lambda = fmin(lambda, x2l);
  if (ang_2 != 0.) lambda = fmin(lambda, x3l);
  const double k_par = TwoPi / lambda;
  const double dby = wave_amplitude * rev[5], dbz = wave_amplitude * rev[6];
  const double bx0 = 1.0, by0 = sqrt(2.0), bz0 = 0.5;

  const int gw = p->gw;
  const long n = orc_ncells(p);
  const double dx = p->dx, dy = p->dy, dz = p->dz;
  double *A = (double *)calloc((size_t)(3 * n), sizeof(double));
  memset(U, 0, sizeof(double) * NV * (size_t)n);
#define AV(i, j, k, c) A[(size_t)(i) + (size_t)p->isize * ((size_t)(j) + (size_t)p->jsize * ((size_t)(k) + (size_t)p->ksize * (size_t)(c)))]
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) { /* :1098-1148 */
        double x = p->xmin + dx / 2 + (i + p->nx * p->px - gw) * dx;
        double y = p->ymin + dy / 2 + (j + p->ny * p->py - gw) * dy;
        double z = p->zmin + dz / 2 + (k + p->nz * p->pz - gw) * dz;
        double Ay, Az, tmpx, tmpy, x1, x2, x3;
        x1 = x; x2 = y - dy / 2; x3 = z - dz / 2;
        tmpx = x1 * cos_a2 * cos_a3 + x2 * cos_a2 * sin_a3 + x3 * sin_a2;
        tmpy = -x1 * sin_a3 + x2 * cos_a3;
        Ay = bz0 * tmpx - (dbz / k_par) * cos(k_par * tmpx);
        Az = -by0 * tmpx + (dby / k_par) * cos(k_par * tmpx) + bx0 * tmpy;
        AV(i, j, k, 0) = -Ay * sin_a3 - Az * sin_a2 * cos_a3;
        x1 = x - dx / 2; x2 = y; x3 = z - dz / 2;
        tmpx = x1 * cos_a2 * cos_a3 + x2 * cos_a2 * sin_a3 + x3 * sin_a2;
        tmpy = -x1 * sin_a3 + x2 * cos_a3;
        Ay = bz0 * tmpx - (dbz / k_par) * cos(k_par * tmpx);
        Az = -by0 * tmpx + (dby / k_par) * cos(k_par * tmpx) + bx0 * tmpy;
        AV(i, j, k, 1) = Ay * cos_a3 - Az * sin_a2 * sin_a3;
        x1 = x - dx / 2; x2 = y - dy / 2; x3 = z;
        tmpx = x1 * cos_a2 * cos_a3 + x2 * cos_a2 * sin_a3 + x3 * sin_a2;
        tmpy = -x1 * sin_a3 + x2 * cos_a3;
        Az = -by0 * tmpx + (dby / k_par) * cos(k_par * tmpx) + bx0 * tmpy;
        AV(i, j, k, 2) = Az * cos_a2;
      }
  for (int k = gw - 1; k < p->ksize - gw + 1; ++k)
    for (int j = gw - 1; j < p->jsize - gw + 1; ++j)
      for (int i = gw - 1; i < p->isize - gw + 1; ++i) { /* :1166-1178 */
        U[AT(p, i, j, k, IA)] = (AV(i, j + 1, k, 2) - AV(i, j, k, 2)) / dy - (AV(i, j, k + 1, 1) - AV(i, j, k, 1)) / dz;
        U[AT(p, i, j, k, IB)] = (AV(i, j, k + 1, 0) - AV(i, j, k, 0)) / dz - (AV(i + 1, j, k, 2) - AV(i, j, k, 2)) / dx;
        U[AT(p, i, j, k, IC)] = (AV(i + 1, j, k, 1) - AV(i, j, k, 1)) / dx - (AV(i, j + 1, k, 0) - AV(i, j, k, 0)) / dy;
      }
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) { /* :1236-1262 */
        double x = p->xmin + dx / 2 + (i + p->nx * p->px - gw) * dx;
        double y = p->ymin + dy / 2 + (j + p->ny * p->py - gw) * dy;
        double z = p->zmin + dz / 2 + (k + p->nz * p->pz - gw) * dz;
        double X = cos_a2 * (x * cos_a3 + y * sin_a3) + z * sin_a2;
        double sn = sin(k_par * X);
        double Mx = d0 * wave_V0 + wave_amplitude * sn * rev[1];
        double My = wave_amplitude * sn * rev[2];
        double Mz = wave_amplitude * sn * rev[3];
        U[AT(p, i, j, k, ID)] = d0 + wave_amplitude * sn * rev[0];
        U[AT(p, i, j, k, IU)] = Mx * cos_a2 * cos_a3 - My * sin_a3 - Mz * sin_a2 * cos_a3;
        U[AT(p, i, j, k, IV)] = Mx * cos_a2 * sin_a3 + My * cos_a3 - Mz * sin_a2 * sin_a3;
        U[AT(p, i, j, k, IW)] = Mx * sin_a2 + Mz * cos_a2;
        U[AT(p, i, j, k, IP)] = p0 / (p->gamma0 - 1.0) + 0.5 * d0 * wave_V0 * wave_V0 + 0.5 * (bx0 * bx0 + by0 * by0 + bz0 * bz0) +
                                wave_amplitude * sn * rev[4];
      }
#undef AV
  free(A);
  return 0;
}

void orc_init_field_loop(const orc_params *p, double radius, double density_in, double amplitude, double vflow,
                         double *U)
{
  /* src/muscl/MHDInitFunctors3D.h:759-1023 : vector potential (only A_z != 0), then interior cells,
   * then interior energy.  Ghost cells stay zero (filled later by make_boundaries). */
  const int gw = p->gw;
  const long n = orc_ncells(p);
  double *Az = (double *)calloc((size_t)n, sizeof(double));
  memset(U, 0, sizeof(double) * NV * (size_t)n);
#define A3(i, j, k) Az[(size_t)(i) + (size_t)p->isize * ((size_t)(j) + (size_t)p->jsize * (size_t)(k))]
  for (int k = 0; k < p->ksize; ++k)
    for (int j = 0; j < p->jsize; ++j)
      for (int i = 0; i < p->isize; ++i) { /* :834-871 */
        double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double r = sqrt(x * x + y * y);
        A3(i, j, k) = 0.0;
        if (r < radius) A3(i, j, k) = amplitude * (radius - r);
      }
  const double cos_theta = 2.0 / sqrt(5.0);
  const double sin_theta = sqrt(1 - cos_theta * cos_theta);
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) { /* :873-975 ; A_x = A_y = 0 */
        double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
        double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
        double r = sqrt(x * x + y * y);
        double d = (r < radius) ? density_in : 1.0;
        U[AT(p, i, j, k, ID)] = d;
        U[AT(p, i, j, k, IU)] = d * vflow * cos_theta;
        U[AT(p, i, j, k, IV)] = d * vflow * sin_theta;
        U[AT(p, i, j, k, IW)] = d * vflow;
        U[AT(p, i, j, k, IA)] = (A3(i, j + 1, k) - A3(i, j, k)) / p->dy - (0.0 - 0.0) / p->dz;
        U[AT(p, i, j, k, IB)] = (0.0 - 0.0) / p->dz - (A3(i + 1, j, k) - A3(i, j, k)) / p->dx;
        U[AT(p, i, j, k, IC)] = (0.0 - 0.0) / p->dx - (0.0 - 0.0) / p->dy;
      }
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) { /* :977-1012 ; note 1.0f and the (still zero) upper ghosts */
        double sa = U[AT(p, i, j, k, IA)] + U[AT(p, i + 1, j, k, IA)];
        double sb = U[AT(p, i, j, k, IB)] + U[AT(p, i, j + 1, k, IB)];
        double sc = U[AT(p, i, j, k, IC)] + U[AT(p, i, j, k + 1, IC)];
        double mu = U[AT(p, i, j, k, IU)], mv = U[AT(p, i, j, k, IV)], mw = U[AT(p, i, j, k, IW)];
        U[AT(p, i, j, k, IP)] = 1.0f / (p->gamma0 - 1.0) +
                                0.5 * (0.25 * (sa * sa) + 0.25 * (sb * sb) + 0.25 * (sc * sc)) +
                                0.5 * (mu * mu + mv * mv + mw * mw) / U[AT(p, i, j, k, ID)];
      }
#undef A3
  free(Az);
}

/* ------------------------------------------------------------------------------------------ */
/* Ghost cells                                                                                */
/* ------------------------------------------------------------------------------------------ */
void orc_make_boundary(const orc_params *p, double *U, int face)
{
  /* src/shared/BoundariesFunctors.h:749-1053 : Dirichlet = mirror with sign flip of the normal
   * momentum and normal B, Neumann = copy of the first interior cell, periodic = wrap.
   * ORC_BC_COPY (interior face of a decomposed run) is filled by the halo exchange instead
   * (SolverBase.cpp:618-691). */
  const int gw = p->gw, nx = p->nx, ny = p->ny, nz = p->nz;
  const int bc = p->bc[face];
  if (bc == ORC_BC_COPY) return;
  const int dir = face / 2, hi = face & 1;
  const int n = dir == 0 ? nx : (dir == 1 ? ny : nz);
  const int vflip = dir == 0 ? IU : (dir == 1 ? IV : IW);
  const int bflip = dir == 0 ? IA : (dir == 1 ? IB : IC);
  /* transverse extents cover the whole array (imin..imax etc.), which is what makes the X,Y,Z
   * sequence fill edges and corners */
  const int e0 = dir == 0 ? p->jsize : p->isize;
  const int e1 = dir == 2 ? p->jsize : p->ksize;
  for (int b = 0; b < e1; ++b)
    for (int a = 0; a < e0; ++a)
      for (int g = 0; g < gw; ++g) {
        int c = hi ? g + n + gw : g; /* ghost index along dir */
        int c0;
        if (bc == ORC_BC_DIRICHLET) c0 = hi ? 2 * n + 2 * gw - 1 - c : 2 * gw - 1 - c;
        else if (bc == ORC_BC_NEUMANN) c0 = hi ? n + gw - 1 : gw;
        else c0 = hi ? c - n : n + c; /* periodic (also what the reference does for any other value) */
        for (int v = 0; v < NV; ++v) {
          double sign = 1.0;
          if (bc == ORC_BC_DIRICHLET && (v == vflip || v == bflip)) sign = -1.0;
          size_t dst, src;
          if (dir == 0) { dst = AT(p, c, a, b, v); src = AT(p, c0, a, b, v); }
          else if (dir == 1) { dst = AT(p, a, c, b, v); src = AT(p, a, c0, b, v); }
          else { dst = AT(p, a, b, c, v); src = AT(p, a, b, c0, v); }
          U[dst] = U[src] * sign;
        }
      }
}

void orc_make_boundaries(const orc_params *p, double *U)
{
  /* src/shared/SolverBase.cpp:527-537 : strictly XMIN,XMAX,YMIN,YMAX,ZMIN,ZMAX */
  for (int f = 0; f < 6; ++f) orc_make_boundary(p, U, f);
}

/* ------------------------------------------------------------------------------------------ */
/* Primitive variables and time step                                                          */
/* ------------------------------------------------------------------------------------------ */
static void constoprim(const orc_params *p, const state_t u, const double bnext[3], state_t q)
{
  /* src/muscl/MHDBaseFunctor3D.h:192-242 (cIso == 0 branch; the sound speed output is unused) */
  q[ID] = fmax(u[ID], p->smallr);
  q[IU] = u[IU] / q[ID];
  q[IV] = u[IV] / q[ID];
  q[IW] = u[IW] / q[ID];
  q[IA] = 0.5 * (u[IA] + bnext[0]);
  q[IB] = 0.5 * (u[IB] + bnext[1]);
  q[IC] = 0.5 * (u[IC] + bnext[2]);
  double eken = 0.5 * (q[IU] * q[IU] + q[IV] * q[IV] + q[IW] * q[IW]);
  double emag = 0.5 * (q[IA] * q[IA] + q[IB] * q[IB] + q[IC] * q[IC]);
  double eint = (u[IP] - emag) / q[ID] - eken;
  q[IP] = fmax((p->gamma0 - 1.0) * q[ID] * eint, q[ID] * p->smallp);
}

void orc_convert_to_primitives(const orc_params *p, const double *U, double *Q)
{
  /* src/muscl/MHDRunFunctors3D.h:107-158 : range [0, size-1) on each axis */
#pragma omp parallel for collapse(2)
  for (int k = 0; k < p->ksize - 1; ++k)
    for (int j = 0; j < p->jsize - 1; ++j)
      for (int i = 0; i < p->isize - 1; ++i) {
        state_t u, q;
        double bn[3];
        load_state(p, U, i, j, k, u);
        bn[0] = U[AT(p, i + 1, j, k, IA)];
        bn[1] = U[AT(p, i, j + 1, k, IB)];
        bn[2] = U[AT(p, i, j, k + 1, IC)];
        constoprim(p, u, bn, q);
        store_state(p, Q, i, j, k, q);
      }
}

static double fast_speed(const orc_params *p, const state_t q, int dir)
{
  /* src/shared/mhd_utils.h:89-117 (find_speed_fast<dir>) */
  double d = q[ID], pr = q[IP], a = q[IA], b = q[IB], c = q[IC];
  double b2 = a * a + b * b + c * c;
  double c2 = p->gamma0 * pr / d;
  double d2 = 0.5 * (b2 / d + c2);
  double n = dir == 0 ? a : (dir == 1 ? b : c);
  return sqrt(d2 + sqrt(d2 * d2 - c2 * n * n / d));
}

double orc_compute_inv_dt(const orc_params *p, const double *Q)
{
  /* src/muscl/MHDRunFunctors3D.h:37-79 with find_speed_info<THREE_D> (mhd_utils.h:319-366) */
  const int gw = p->gw;
  double invDt = 0.0; /* SolverMHDMuscl.h:729 */
#pragma omp parallel for collapse(2) reduction(max : invDt)
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) {
        state_t q;
        load_state(p, Q, i, j, k, q);
        double vx = fast_speed(p, q, 0) + fabs(q[IU]);
        double vy = fast_speed(p, q, 1) + fabs(q[IV]);
        double vz = fast_speed(p, q, 2) + fabs(q[IW]);
        invDt = fmax(invDt, vx / p->dx + vy / p->dy + vz / p->dz);
      }
  return invDt;
}

double orc_compute_dt_local(const orc_params *p, const double *Q)
{
  return p->cfl / orc_compute_inv_dt(p, Q); /* SolverMHDMuscl.h:737 */
}

/* ------------------------------------------------------------------------------------------ */
/* v0 scratch                                                                                 */
/* ------------------------------------------------------------------------------------------ */
enum { S_QM_X, S_QM_Y, S_QM_Z, S_QP_X, S_QP_Y, S_QP_Z,
       S_RT, S_RB, S_LT, S_LB, S_RT2, S_RB2, S_LT2, S_LB2, S_RT3, S_RB3, S_LT3, S_LB3,
       S_FX, S_FY, S_FZ, S_N8 };
static const char *const s_names8[S_N8] = { "Qm_x", "Qm_y", "Qm_z", "Qp_x", "Qp_y", "Qp_z",
  "QEdge_RT", "QEdge_RB", "QEdge_LT", "QEdge_LB", "QEdge_RT2", "QEdge_RB2", "QEdge_LT2", "QEdge_LB2",
  "QEdge_RT3", "QEdge_RB3", "QEdge_LT3", "QEdge_LB3", "Fluxes_x", "Fluxes_y", "Fluxes_z" };
enum { S_ELEC, S_DA, S_DB, S_DC, S_EMF, S_N3 };
static const char *const s_names3[S_N3] = { "ElecField", "DeltaA", "DeltaB", "DeltaC", "Emf" };

struct orc_scratch {
  double *a8[S_N8]; /* 8-component arrays (SolverMHDMuscl.h:326-386) */
  double *a3[S_N3]; /* 3-component arrays */
};

orc_scratch *orc_scratch_create(const orc_params *p)
{
  orc_scratch *s = (orc_scratch *)calloc(1, sizeof(*s));
  size_t n = (size_t)orc_ncells(p);
  for (int a = 0; a < S_N8; ++a) s->a8[a] = (double *)calloc(n * NV, sizeof(double));
  for (int a = 0; a < S_N3; ++a) s->a3[a] = (double *)calloc(n * 3, sizeof(double));
  return s;
}
void orc_scratch_destroy(orc_scratch *s)
{
  if (!s) return;
  for (int a = 0; a < S_N8; ++a) free(s->a8[a]);
  for (int a = 0; a < S_N3; ++a) free(s->a3[a]);
  free(s);
}
const double *orc_scratch_array(const orc_scratch *s, const char *name)
{
  for (int a = 0; a < S_N8; ++a) if (!strcmp(name, s_names8[a])) return s->a8[a];
  for (int a = 0; a < S_N3; ++a) if (!strcmp(name, s_names3[a])) return s->a3[a];
  return NULL;
}

/* ------------------------------------------------------------------------------------------ */
/* Edge electric field and face-B slopes                                                      */
/* ------------------------------------------------------------------------------------------ */
static void compute_elec_field(const orc_params *p, const double *U, const double *Q, double *E)
{
  /* src/muscl/MHDRunFunctors3D.h:301-356 : range [1, size-1) ; 4-cell average in this exact order */
#pragma omp parallel for collapse(2)
  for (int k = 1; k < p->ksize - 1; ++k)
    for (int j = 1; j < p->jsize - 1; ++j)
      for (int i = 1; i < p->isize - 1; ++i) {
        double u, v, w, A, B, C;
        v = 0.25 * (Q[AT(p, i, j - 1, k - 1, IV)] + Q[AT(p, i, j - 1, k, IV)] + Q[AT(p, i, j, k - 1, IV)] + Q[AT(p, i, j, k, IV)]);
        w = 0.25 * (Q[AT(p, i, j - 1, k - 1, IW)] + Q[AT(p, i, j - 1, k, IW)] + Q[AT(p, i, j, k - 1, IW)] + Q[AT(p, i, j, k, IW)]);
        B = 0.5 * (U[AT(p, i, j, k - 1, IB)] + U[AT(p, i, j, k, IB)]);
        C = 0.5 * (U[AT(p, i, j - 1, k, IC)] + U[AT(p, i, j, k, IC)]);
        E[AT(p, i, j, k, 0)] = v * C - w * B;

        u = 0.25 * (Q[AT(p, i - 1, j, k - 1, IU)] + Q[AT(p, i - 1, j, k, IU)] + Q[AT(p, i, j, k - 1, IU)] + Q[AT(p, i, j, k, IU)]);
        w = 0.25 * (Q[AT(p, i - 1, j, k - 1, IW)] + Q[AT(p, i - 1, j, k, IW)] + Q[AT(p, i, j, k - 1, IW)] + Q[AT(p, i, j, k, IW)]);
        A = 0.5 * (U[AT(p, i, j, k - 1, IA)] + U[AT(p, i, j, k, IA)]);
        C = 0.5 * (U[AT(p, i - 1, j, k, IC)] + U[AT(p, i, j, k, IC)]);
        E[AT(p, i, j, k, 1)] = w * A - u * C;

        u = 0.25 * (Q[AT(p, i - 1, j - 1, k, IU)] + Q[AT(p, i - 1, j, k, IU)] + Q[AT(p, i, j - 1, k, IU)] + Q[AT(p, i, j, k, IU)]);
        v = 0.25 * (Q[AT(p, i - 1, j - 1, k, IV)] + Q[AT(p, i - 1, j, k, IV)] + Q[AT(p, i, j - 1, k, IV)] + Q[AT(p, i, j, k, IV)]);
        A = 0.5 * (U[AT(p, i, j - 1, k, IA)] + U[AT(p, i, j, k, IA)]);
        B = 0.5 * (U[AT(p, i - 1, j, k, IB)] + U[AT(p, i, j, k, IB)]);
        E[AT(p, i, j, k, 2)] = u * B - v * A;
      }
}

static double limited_slope(double st, double q, double qplus, double qminus)
{
  /* src/muscl/MHDBaseFunctor3D.h:280-288 (same code at :605-665 for face B) */
  double dlft = st * (q - qminus);
  double drgt = st * (qplus - q);
  double dcen = 0.5 * (qplus - qminus);
  double dsgn = (dcen >= 0.0) ? 1.0 : -1.0;
  double slop = fmin(fabs(dlft), fabs(drgt));
  double dlim = slop;
  if ((dlft * drgt) <= 0.0) dlim = 0.0;
  return dsgn * fmin(dlim, fabs(dcen));
}

static void compute_mag_slopes(const orc_params *p, const double *U, double *dA, double *dB, double *dC)
{
  /* src/muscl/MHDRunFunctors3D.h:470-532 + slope_unsplit_mhd_3d (MHDBaseFunctor3D.h:561-668):
   * Delta?(.., 0|1|2) = slope along x|y|z ; the slope of a component along its own normal stays 0. */
  const double st = fmin(p->slope_type, 2.0);
#pragma omp parallel for collapse(2)
  for (int k = 1; k < p->ksize - 1; ++k)
    for (int j = 1; j < p->jsize - 1; ++j)
      for (int i = 1; i < p->isize - 1; ++i) {
        double a = U[AT(p, i, j, k, IA)], b = U[AT(p, i, j, k, IB)], c = U[AT(p, i, j, k, IC)];
        dA[AT(p, i, j, k, 0)] = 0.0;
        dA[AT(p, i, j, k, 1)] = limited_slope(st, a, U[AT(p, i, j + 1, k, IA)], U[AT(p, i, j - 1, k, IA)]);
        dA[AT(p, i, j, k, 2)] = limited_slope(st, a, U[AT(p, i, j, k + 1, IA)], U[AT(p, i, j, k - 1, IA)]);
        dB[AT(p, i, j, k, 0)] = limited_slope(st, b, U[AT(p, i + 1, j, k, IB)], U[AT(p, i - 1, j, k, IB)]);
        dB[AT(p, i, j, k, 1)] = 0.0;
        dB[AT(p, i, j, k, 2)] = limited_slope(st, b, U[AT(p, i, j, k + 1, IB)], U[AT(p, i, j, k - 1, IB)]);
        dC[AT(p, i, j, k, 0)] = limited_slope(st, c, U[AT(p, i + 1, j, k, IC)], U[AT(p, i - 1, j, k, IC)]);
        dC[AT(p, i, j, k, 1)] = limited_slope(st, c, U[AT(p, i, j + 1, k, IC)], U[AT(p, i, j - 1, k, IC)]);
        dC[AT(p, i, j, k, 2)] = 0.0;
      }
}

/* ------------------------------------------------------------------------------------------ */
/* Trace (Hancock predictor): 6 face states + 12 edge states per cell                         */
/* ------------------------------------------------------------------------------------------ */
static void hydro_slopes(const orc_params *p, const state_t q, const state_t qpx, const state_t qmx,
                         const state_t qpy, const state_t qmy, const state_t qpz, const state_t qmz,
                         state_t dq[3])
{
  /* src/muscl/MHDBaseFunctor3D.h:362-495 */
  const double st = p->slope_type;
  if (st == 1 || st == 2) {
    for (int v = 0; v < NV; ++v) {
      dq[0][v] = limited_slope(st, q[v], qpx[v], qmx[v]);
      dq[1][v] = limited_slope(st, q[v], qpy[v], qmy[v]);
      dq[2][v] = limited_slope(st, q[v], qpz[v], qmz[v]);
    }
    /* :452-461 passes qMinusY[IC] as the z-minus neighbour of the last component */
    dq[2][IC] = limited_slope(st, q[IC], qpz[IC], qmy[IC]);
  } else {
    for (int d = 0; d < 3; ++d)
      for (int v = 0; v < NV; ++v) dq[d][v] = 0.0;
  }
}

static void floor_state(const orc_params *p, state_t s)
{
  /* MHDBaseFunctor3D.h:907-908 etc.: density floor smallr, pressure floor smallp (NOT times rho) */
  s[ID] = fmax(p->smallr, s[ID]);
  s[IP] = fmax(p->smallp, s[IP]);
}

static void trace_cell(const orc_params *p, const state_t q, state_t dq[3], const double bf[6],
                       const double dbf[12], const double E[3][2][2], double dtdx, double dtdy, double dtdz,
                       state_t qm[3], state_t qp[3], state_t qe[4][3])
{
  /* src/muscl/MHDBaseFunctor3D.h:688-1114 (trace_unsplit_mhd_3d_simpler), Omega0 == 0 */
  const double gamma = p->gamma0;
  const double ELL = E[0][0][0], ELR = E[0][0][1], ERL = E[0][1][0], ERR = E[0][1][1];
  const double FLL = E[1][0][0], FLR = E[1][0][1], FRL = E[1][1][0], FRR = E[1][1][1];
  const double GLL = E[2][0][0], GLR = E[2][0][1], GRL = E[2][1][0], GRR = E[2][1][1];

  double r = q[ID], pr = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  double AL = bf[0], AR = bf[1], BL = bf[2], BR = bf[3], CL = bf[4], CR = bf[5];

  /* :767-813 slopes are halved in place */
  double drx = (dq[0][ID] *= 0.5), dpx = (dq[0][IP] *= 0.5), dux = (dq[0][IU] *= 0.5), dvx = (dq[0][IV] *= 0.5),
         dwx = (dq[0][IW] *= 0.5), dCx = (dq[0][IC] *= 0.5), dBx = (dq[0][IB] *= 0.5);
  double dry = (dq[1][ID] *= 0.5), dpy = (dq[1][IP] *= 0.5), duy = (dq[1][IU] *= 0.5), dvy = (dq[1][IV] *= 0.5),
         dwy = (dq[1][IW] *= 0.5), dCy = (dq[1][IC] *= 0.5), dAy = (dq[1][IA] *= 0.5);
  double drz = (dq[2][ID] *= 0.5), dpz = (dq[2][IP] *= 0.5), duz = (dq[2][IU] *= 0.5), dvz = (dq[2][IV] *= 0.5),
         dwz = (dq[2][IW] *= 0.5), dAz = (dq[2][IA] *= 0.5), dBz = (dq[2][IB] *= 0.5);

  /* :816-829 */
  double dALy = 0.5 * dbf[0], dALz = 0.5 * dbf[1], dBLx = 0.5 * dbf[2], dBLz = 0.5 * dbf[3], dCLx = 0.5 * dbf[4],
         dCLy = 0.5 * dbf[5];
  double dARy = 0.5 * dbf[6], dARz = 0.5 * dbf[7], dBRx = 0.5 * dbf[8], dBRz = 0.5 * dbf[9], dCRx = 0.5 * dbf[10],
         dCRy = 0.5 * dbf[11];
  /* :832-834 */
  double dAx = 0.5 * (AR - AL), dBy = 0.5 * (BR - BL), dCz = 0.5 * (CR - CL);

  /* :843-857 source terms */
  double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
  double su0 = (-u * dux - (dpx + B * dBx + C * dCx) / r) * dtdx + (-v * duy + B * dAy / r) * dtdy +
               (-w * duz + C * dAz / r) * dtdz;
  double sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) / r) * dtdy +
               (-w * dvz + C * dBz / r) * dtdz;
  double sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy +
               (-w * dwz - (dpz + A * dAz + B * dBz) / r) * dtdz;
  double sp0 = (-u * dpx - dux * gamma * pr) * dtdx + (-v * dpy - dvy * gamma * pr) * dtdy +
               (-w * dpz - dwz * gamma * pr) * dtdz;
  double sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy + (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
  double sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx + (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
  double sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;

  /* :872-877 face-centred B from the edge electric fields */
  double sAL0 = +(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5;
  double sAR0 = +(GRR - GRL) * dtdy * 0.5 - (FRR - FRL) * dtdz * 0.5;
  double sBL0 = -(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5;
  double sBR0 = -(GRR - GLR) * dtdx * 0.5 + (ERR - ERL) * dtdz * 0.5;
  double sCL0 = +(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5;
  double sCR0 = +(FRR - FLR) * dtdx * 0.5 - (ERR - ELR) * dtdy * 0.5;

  /* :882-896 */
  r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; pr = pr + sp0; A = A + sA0; B = B + sB0; C = C + sC0;
  AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0; CL = CL + sCL0; CR = CR + sCR0;

#define SET(S, R_, U_, V_, W_, P_, A_, B_, C_) \
  do { (S)[ID] = (R_); (S)[IU] = (U_); (S)[IV] = (V_); (S)[IW] = (W_); (S)[IP] = (P_); \
       (S)[IA] = (A_); (S)[IB] = (B_); (S)[IC] = (C_); floor_state(p, (S)); } while (0)

  /* :898-968 face states */
  SET(qp[0], r - drx, u - dux, v - dvx, w - dwx, pr - dpx, AL, B - dBx, C - dCx);
  SET(qm[0], r + drx, u + dux, v + dvx, w + dwx, pr + dpx, AR, B + dBx, C + dCx);
  SET(qp[1], r - dry, u - duy, v - dvy, w - dwy, pr - dpy, A - dAy, BL, C - dCy);
  SET(qm[1], r + dry, u + duy, v + dvy, w + dwy, pr + dpy, A + dAy, BR, C + dCy);
  SET(qp[2], r - drz, u - duz, v - dvz, w - dwz, pr - dpz, A - dAz, B - dBz, CL);
  SET(qm[2], r + drz, u + duz, v + dvz, w + dwz, pr + dpz, A + dAz, B + dBz, CR);

  /* :970-1016 X-edges */
  SET(qe[IRT][0], r + (+dry + drz), u + (+duy + duz), v + (+dvy + dvz), w + (+dwy + dwz), pr + (+dpy + dpz),
      A + (+dAy + dAz), BR + (+dBRz), CR + (+dCRy));
  SET(qe[IRB][0], r + (+dry - drz), u + (+duy - duz), v + (+dvy - dvz), w + (+dwy - dwz), pr + (+dpy - dpz),
      A + (+dAy - dAz), BR + (-dBRz), CL + (+dCLy));
  SET(qe[ILT][0], r + (-dry + drz), u + (-duy + duz), v + (-dvy + dvz), w + (-dwy + dwz), pr + (-dpy + dpz),
      A + (-dAy + dAz), BL + (+dBLz), CR + (-dCRy));
  SET(qe[ILB][0], r + (-dry - drz), u + (-duy - duz), v + (-dvy - dvz), w + (-dwy - dwz), pr + (-dpy - dpz),
      A + (-dAy - dAz), BL + (-dBLz), CL + (-dCLy));
  /* :1018-1064 Y-edges */
  SET(qe[IRT][1], r + (+drx + drz), u + (+dux + duz), v + (+dvx + dvz), w + (+dwx + dwz), pr + (+dpx + dpz),
      AR + (+dARz), B + (+dBx + dBz), CR + (+dCRx));
  SET(qe[IRB][1], r + (+drx - drz), u + (+dux - duz), v + (+dvx - dvz), w + (+dwx - dwz), pr + (+dpx - dpz),
      AR + (-dARz), B + (+dBx - dBz), CL + (+dCLx));
  SET(qe[ILT][1], r + (-drx + drz), u + (-dux + duz), v + (-dvx + dvz), w + (-dwx + dwz), pr + (-dpx + dpz),
      AL + (+dALz), B + (-dBx + dBz), CR + (-dCRx));
  SET(qe[ILB][1], r + (-drx - drz), u + (-dux - duz), v + (-dvx - dvz), w + (-dwx - dwz), pr + (-dpx - dpz),
      AL + (-dALz), B + (-dBx - dBz), CL + (-dCLx));
  /* :1066-1112 Z-edges */
  SET(qe[IRT][2], r + (+drx + dry), u + (+dux + duy), v + (+dvx + dvy), w + (+dwx + dwy), pr + (+dpx + dpy),
      AR + (+dARy), BR + (+dBRx), C + (+dCx + dCy));
  SET(qe[IRB][2], r + (+drx - dry), u + (+dux - duy), v + (+dvx - dvy), w + (+dwx - dwy), pr + (+dpx - dpy),
      AR + (-dARy), BL + (+dBLx), C + (+dCx - dCy));
  SET(qe[ILT][2], r + (-drx + dry), u + (-dux + duy), v + (-dvx + dvy), w + (-dwx + dwy), pr + (-dpx + dpy),
      AL + (+dALy), BR + (-dBRx), C + (-dCx + dCy));
  SET(qe[ILB][2], r + (-drx - dry), u + (-dux - duy), v + (-dvx - dvy), w + (-dwx - dwy), pr + (-dpx - dpy),
      AL + (-dALy), BL + (-dBLx), C + (-dCx - dCy));
#undef SET
}

static void compute_trace(const orc_params *p, const double *U, const double *Q, orc_scratch *s, double dtdx,
                          double dtdy, double dtdz)
{
  /* src/muscl/MHDRunFunctors3D.h:668-844 : range [gw-2, size-gw+1) */
  const int gw = p->gw;
  const double *E = s->a3[S_ELEC], *dA = s->a3[S_DA], *dB = s->a3[S_DB], *dC = s->a3[S_DC];
#pragma omp parallel for collapse(2)
  for (int k = gw - 2; k < p->ksize - gw + 1; ++k)
    for (int j = gw - 2; j < p->jsize - gw + 1; ++j)
      for (int i = gw - 2; i < p->isize - gw + 1; ++i) {
        state_t q, qpx, qmx, qpy, qmy, qpz, qmz, dq[3], qm[3], qp[3], qe[4][3];
        double bf[6], dbf[12], el[3][2][2];
        load_state(p, Q, i, j, k, q);
        load_state(p, Q, i + 1, j, k, qpx);
        load_state(p, Q, i - 1, j, k, qmx);
        load_state(p, Q, i, j + 1, k, qpy);
        load_state(p, Q, i, j - 1, k, qmy);
        load_state(p, Q, i, j, k + 1, qpz);
        load_state(p, Q, i, j, k - 1, qmz);
        hydro_slopes(p, q, qpx, qmx, qpy, qmy, qpz, qmz, dq);

        bf[0] = U[AT(p, i, j, k, IA)]; bf[1] = U[AT(p, i + 1, j, k, IA)];
        bf[2] = U[AT(p, i, j, k, IB)]; bf[3] = U[AT(p, i, j + 1, k, IB)];
        bf[4] = U[AT(p, i, j, k, IC)]; bf[5] = U[AT(p, i, j, k + 1, IC)];

        dbf[0] = dA[AT(p, i, j, k, 1)]; dbf[1] = dA[AT(p, i, j, k, 2)];
        dbf[2] = dB[AT(p, i, j, k, 0)]; dbf[3] = dB[AT(p, i, j, k, 2)];
        dbf[4] = dC[AT(p, i, j, k, 0)]; dbf[5] = dC[AT(p, i, j, k, 1)];
        dbf[6] = dA[AT(p, i + 1, j, k, 1)]; dbf[7] = dA[AT(p, i + 1, j, k, 2)];
        dbf[8] = dB[AT(p, i, j + 1, k, 0)]; dbf[9] = dB[AT(p, i, j + 1, k, 2)];
        dbf[10] = dC[AT(p, i, j, k + 1, 0)]; dbf[11] = dC[AT(p, i, j, k + 1, 1)];

        el[0][0][0] = E[AT(p, i, j, k, 0)]; el[0][0][1] = E[AT(p, i, j, k + 1, 0)];
        el[0][1][0] = E[AT(p, i, j + 1, k, 0)]; el[0][1][1] = E[AT(p, i, j + 1, k + 1, 0)];
        el[1][0][0] = E[AT(p, i, j, k, 1)]; el[1][0][1] = E[AT(p, i, j, k + 1, 1)];
        el[1][1][0] = E[AT(p, i + 1, j, k, 1)]; el[1][1][1] = E[AT(p, i + 1, j, k + 1, 1)];
        el[2][0][0] = E[AT(p, i, j, k, 2)]; el[2][0][1] = E[AT(p, i, j + 1, k, 2)];
        el[2][1][0] = E[AT(p, i + 1, j, k, 2)]; el[2][1][1] = E[AT(p, i + 1, j + 1, k, 2)];

        trace_cell(p, q, dq, bf, dbf, el, dtdx, dtdy, dtdz, qm, qp, qe);

        store_state(p, s->a8[S_QM_X], i, j, k, qm[0]); store_state(p, s->a8[S_QP_X], i, j, k, qp[0]);
        store_state(p, s->a8[S_QM_Y], i, j, k, qm[1]); store_state(p, s->a8[S_QP_Y], i, j, k, qp[1]);
        store_state(p, s->a8[S_QM_Z], i, j, k, qm[2]); store_state(p, s->a8[S_QP_Z], i, j, k, qp[2]);
        store_state(p, s->a8[S_RT], i, j, k, qe[IRT][0]); store_state(p, s->a8[S_RB], i, j, k, qe[IRB][0]);
        store_state(p, s->a8[S_LT], i, j, k, qe[ILT][0]); store_state(p, s->a8[S_LB], i, j, k, qe[ILB][0]);
        store_state(p, s->a8[S_RT2], i, j, k, qe[IRT][1]); store_state(p, s->a8[S_RB2], i, j, k, qe[IRB][1]);
        store_state(p, s->a8[S_LT2], i, j, k, qe[ILT][1]); store_state(p, s->a8[S_LB2], i, j, k, qe[ILB][1]);
        store_state(p, s->a8[S_RT3], i, j, k, qe[IRT][2]); store_state(p, s->a8[S_RB3], i, j, k, qe[IRB][2]);
        store_state(p, s->a8[S_LT3], i, j, k, qe[ILT][2]); store_state(p, s->a8[S_LB3], i, j, k, qe[ILB][2]);
      }
}

/* ------------------------------------------------------------------------------------------ */
/* HLLD face flux                                                                             */
/* ------------------------------------------------------------------------------------------ */
static void riemann_hlld(const orc_params *p, state_t ql, state_t qr, state_t flux)
{
  /* src/shared/RiemannSolvers_MHD.h:133-367 (cIso == 0).  Mutates ql[IA], qr[IA] like the reference. */
  const double entho = 1.0 / (p->gamma0 - 1.0);
  double a = 0.5 * (ql[IA] + qr[IA]);
  double sgnm = (a >= 0) ? 1.0 : -1.0;
  ql[IA] = a;
  qr[IA] = a;

  double rl = ql[ID], pl = ql[IP], ul = ql[IU], vl = ql[IV], wl = ql[IW], bl = ql[IB], cl = ql[IC];
  double ecinl = 0.5 * (ul * ul + vl * vl + wl * wl) * rl;
  double emagl = 0.5 * (a * a + bl * bl + cl * cl);
  double etotl = pl * entho + ecinl + emagl;
  double ptotl = pl + emagl;
  double vdotbl = ul * a + vl * bl + wl * cl;

  double rr = qr[ID], pr = qr[IP], ur = qr[IU], vr = qr[IV], wr = qr[IW], br = qr[IB], cr = qr[IC];
  double ecinr = 0.5 * (ur * ur + vr * vr + wr * wr) * rr;
  double emagr = 0.5 * (a * a + br * br + cr * cr);
  double etotr = pr * entho + ecinr + emagr;
  double ptotr = pr + emagr;
  double vdotbr = ur * a + vr * br + wr * cr;

  double cfastl = fast_speed(p, ql, 0);
  double cfastr = fast_speed(p, qr, 0);

  double sl = fmin(ul, ur) - fmax(cfastl, cfastr);
  double sr = fmax(ul, ur) + fmax(cfastl, cfastr);

  double rcl = rl * (ul - sl);
  double rcr = rr * (sr - ur);

  double ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) / (rcr + rcl);
  double ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);

  /* left star region :204-233 */
  double rstarl = rl * (sl - ul) / (sl - ustar);
  double estar = rl * (sl - ul) * (sl - ustar) - a * a;
  double el = rl * (sl - ul) * (sl - ul) - a * a;
  double vstarl, wstarl, bstarl, cstarl;
  if (a * a > 0 && fabs(estar / (a * a) - 1.0) <= 1e-8) {
    vstarl = vl; bstarl = bl; wstarl = wl; cstarl = cl;
  } else {
    vstarl = vl - a * bl * (ustar - ul) / estar;
    bstarl = bl * el / estar;
    wstarl = wl - a * cl * (ustar - ul) / estar;
    cstarl = cl * el / estar;
  }
  double vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  double etotstarl = ((sl - ul) * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) / (sl - ustar);
  double sqrrstarl = sqrt(rstarl);
  double calfvenl = fabs(a) / sqrrstarl;
  double sal = ustar - calfvenl;

  /* right star region :235-263 */
  double rstarr = rr * (sr - ur) / (sr - ustar);
  estar = rr * (sr - ur) * (sr - ustar) - a * a;
  double er = rr * (sr - ur) * (sr - ur) - a * a;
  double vstarr, wstarr, bstarr, cstarr;
  if (a * a > 0 && fabs(estar / (a * a) - 1.0) <= 1e-8) {
    vstarr = vr; bstarr = br; wstarr = wr; cstarr = cr;
  } else {
    vstarr = vr - a * br * (ustar - ur) / estar;
    bstarr = br * er / estar;
    wstarr = wr - a * cr * (ustar - ur) / estar;
    cstarr = cr * er / estar;
  }
  double vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  double etotstarr = ((sr - ur) * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) / (sr - ustar);
  double sqrrstarr = sqrt(rstarr);
  double calfvenr = fabs(a) / sqrrstarr;
  double sar = ustar + calfvenr;

  /* double star region :265-278 */
  double vstarstar = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (bstarr - bstarl)) / (sqrrstarl + sqrrstarr);
  double wstarstar = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (cstarr - cstarl)) / (sqrrstarl + sqrrstarr);
  double bstarstar =
    (sqrrstarl * bstarr + sqrrstarr * bstarl + sgnm * sqrrstarl * sqrrstarr * (vstarr - vstarl)) / (sqrrstarl + sqrrstarr);
  double cstarstar =
    (sqrrstarl * cstarr + sqrrstarr * cstarl + sgnm * sqrrstarl * sqrrstarr * (wstarr - wstarl)) / (sqrrstarl + sqrrstarr);
  double vdotbstarstar = ustar * a + vstarstar * bstarstar + wstarstar * cstarstar;
  double etotstarstarl = etotstarl - sgnm * sqrrstarl * (vdotbstarl - vdotbstarstar);
  double etotstarstarr = etotstarr + sgnm * sqrrstarr * (vdotbstarr - vdotbstarstar);

  /* sample at x/t = 0 :280-353 */
  double ro, uo, vo, wo, bo, co, ptoto, etoto, vdotbo;
  if (sl > 0) {
    ro = rl; uo = ul; vo = vl; wo = wl; bo = bl; co = cl; ptoto = ptotl; etoto = etotl; vdotbo = vdotbl;
  } else if (sal > 0) {
    ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; bo = bstarl; co = cstarl; ptoto = ptotstar; etoto = etotstarl; vdotbo = vdotbstarl;
  } else if (ustar > 0) {
    ro = rstarl; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarl; vdotbo = vdotbstarstar;
  } else if (sar > 0) {
    ro = rstarr; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarr; vdotbo = vdotbstarstar;
  } else if (sr > 0) {
    ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; bo = bstarr; co = cstarr; ptoto = ptotstar; etoto = etotstarr; vdotbo = vdotbstarr;
  } else {
    ro = rr; uo = ur; vo = vr; wo = wr; bo = br; co = cr; ptoto = ptotr; etoto = etotr; vdotbo = vdotbr;
  }

  /* :355-365 */
  flux[ID] = ro * uo;
  flux[IP] = (etoto + ptoto) * uo - a * vdotbo;
  flux[IU] = ro * uo * uo - a * a + ptoto;
  flux[IV] = ro * uo * vo - a * bo;
  flux[IW] = ro * uo * wo - a * co;
  flux[IA] = 0.0;
  flux[IB] = bo * uo - a * vo;
  flux[IC] = co * uo - a * wo;
}

static void find_mhd_flux(const orc_params *p, const state_t q, state_t cvar, state_t ff)
{
  /* src/shared/mhd_utils.h:175-231 (cIso == 0) */
  const double entho = 1.0 / (p->gamma0 - 1.0);
  double d = q[ID], pr = q[IP], u = q[IU], v = q[IV], w = q[IW], a = q[IA], b = q[IB], c = q[IC];
  double ecin = 0.5 * (u * u + v * v + w * w) * d;
  double emag = 0.5 * (a * a + b * b + c * c);
  double etot = pr * entho + ecin + emag;
  double ptot = pr + emag;
  cvar[ID] = d; cvar[IP] = etot; cvar[IU] = d * u; cvar[IV] = d * v; cvar[IW] = d * w;
  cvar[IA] = a; cvar[IB] = b; cvar[IC] = c;
  ff[ID] = d * u;
  ff[IP] = (etot + ptot) * u - a * (a * u + b * v + c * w);
  ff[IU] = d * u * u - a * a + ptot;
  ff[IV] = d * u * v - a * b;
  ff[IW] = d * u * w - a * c;
  ff[IA] = 0.0;
  ff[IB] = b * u - a * v;
  ff[IC] = c * u - a * w;
}

static void riemann_hll(const orc_params *p, state_t ql, state_t qr, state_t flux)
{
  /* src/shared/RiemannSolvers_MHD.h:27-69 */
  double bx_mean = 0.5 * (ql[IA] + qr[IA]);
  ql[IA] = bx_mean;
  qr[IA] = bx_mean;
  state_t ul, fl, ur, fr;
  find_mhd_flux(p, ql, ul, fl);
  find_mhd_flux(p, qr, ur, fr);
  double cfl_ = fast_speed(p, ql, 0), cfr = fast_speed(p, qr, 0);
  double vl = ql[IU], vr = qr[IU];
  double sl = fmin(fmin(vl, vr) - fmax(cfl_, cfr), 0.0);
  double sr = fmax(fmax(vl, vr) + fmax(cfl_, cfr), 0.0);
  for (int v = 0; v < NV; ++v) flux[v] = (sr * fl[v] - sl * fr[v] + sr * sl * (ur[v] - ul[v])) / (sr - sl);
}

static void riemann_llf(const orc_params *p, state_t ql, state_t qr, state_t flux)
{
  /* src/shared/RiemannSolvers_MHD.h:83-111 ; find_speed_info (scalar overload) mhd_utils.h:377-402 */
  double bx_mean = 0.5 * (ql[IA] + qr[IA]);
  ql[IA] = bx_mean;
  qr[IA] = bx_mean;
  state_t ul, fl, ur, fr;
  find_mhd_flux(p, ql, ul, fl);
  find_mhd_flux(p, qr, ur, fr);
  for (int v = 0; v < NV; ++v) flux[v] = (fl[v] + fr[v]) / 2;
  double cleft = fast_speed(p, ql, 0) + fabs(ql[IU]);
  double cright = fast_speed(p, qr, 0) + fabs(qr[IU]);
  double vel_info = fmax(cleft, cright);
  for (int v = 0; v < NV; ++v) flux[v] -= vel_info * (ur[v] - ul[v]) / 2;
}

static void riemann_mhd(const orc_params *p, state_t ql, state_t qr, state_t flux)
{
  /* src/shared/RiemannSolvers_MHD.h:372-392 */
  if (p->riemann == 2) riemann_hll(p, ql, qr, flux);
  else if (p->riemann == 1) riemann_llf(p, ql, qr, flux);
  else riemann_hlld(p, ql, qr, flux);
}

static void swap2(double *a, double *b) { double t = *a; *a = *b; *b = t; }

static void compute_fluxes(const orc_params *p, orc_scratch *s)
{
  /* src/muscl/MHDRunFunctors3D.h:1837-1902 : range [gw, size-gw+1) ; y/z fluxes stay in the rotated frame */
  const int gw = p->gw;
#pragma omp parallel for collapse(2)
  for (int k = gw; k < p->ksize - gw + 1; ++k)
    for (int j = gw; j < p->jsize - gw + 1; ++j)
      for (int i = gw; i < p->isize - gw + 1; ++i) {
        state_t ql, qr, f;
        load_state(p, s->a8[S_QM_X], i - 1, j, k, ql);
        load_state(p, s->a8[S_QP_X], i, j, k, qr);
        riemann_mhd(p, ql, qr, f);
        store_state(p, s->a8[S_FX], i, j, k, f);

        load_state(p, s->a8[S_QM_Y], i, j - 1, k, ql);
        swap2(&ql[IU], &ql[IV]); swap2(&ql[IA], &ql[IB]);
        load_state(p, s->a8[S_QP_Y], i, j, k, qr);
        swap2(&qr[IU], &qr[IV]); swap2(&qr[IA], &qr[IB]);
        riemann_mhd(p, ql, qr, f);
        store_state(p, s->a8[S_FY], i, j, k, f);

        load_state(p, s->a8[S_QM_Z], i, j, k - 1, ql);
        swap2(&ql[IU], &ql[IW]); swap2(&ql[IA], &ql[IC]);
        load_state(p, s->a8[S_QP_Z], i, j, k, qr);
        swap2(&qr[IU], &qr[IW]); swap2(&qr[IA], &qr[IC]);
        riemann_mhd(p, ql, qr, f);
        store_state(p, s->a8[S_FZ], i, j, k, f);
      }
}

/* ------------------------------------------------------------------------------------------ */
/* Edge EMFs: 2-D magnetic HLLD                                                               */
/* ------------------------------------------------------------------------------------------ */
static double max4(double a0, double a1, double a2, double a3)
{ /* mhd_utils.h:36-46 comparison chain */
  double r = a0; r = (a1 > r) ? a1 : r; r = (a2 > r) ? a2 : r; r = (a3 > r) ? a3 : r; return r;
}
static double min4(double a0, double a1, double a2, double a3)
{ /* mhd_utils.h:51-61 */
  double r = a0; r = (a1 < r) ? a1 : r; r = (a2 < r) ? a2 : r; r = (a3 < r) ? a3 : r; return r;
}
static double max5(double a0, double a1, double a2, double a3, double a4)
{ /* mhd_utils.h:66-77 */
  double r = a0; r = (a1 > r) ? a1 : r; r = (a2 > r) ? a2 : r; r = (a3 > r) ? a3 : r; r = (a4 > r) ? a4 : r; return r;
}

static double mag_riemann2d_hlld(const orc_params *p, state_t qLLRR[4], const double eLLRR[4])
{
  /* src/shared/RiemannSolvers_MHD.h:398-630 */
  const double *qLL = qLLRR[ILL], *qRL = qLLRR[IRL], *qLR = qLLRR[ILR], *qRR = qLLRR[IRR];
  const double ELL = eLLRR[ILL], ERL = eLLRR[IRL], ELR = eLLRR[ILR], ERR = eLLRR[IRR];
  const double rLL = qLL[ID], pLL = qLL[IP], uLL = qLL[IU], vLL = qLL[IV], aLL = qLL[IA], bLL = qLL[IB], cLL = qLL[IC];
  const double rLR = qLR[ID], pLR = qLR[IP], uLR = qLR[IU], vLR = qLR[IV], aLR = qLR[IA], bLR = qLR[IB], cLR = qLR[IC];
  const double rRL = qRL[ID], pRL = qRL[IP], uRL = qRL[IU], vRL = qRL[IV], aRL = qRL[IA], bRL = qRL[IB], cRL = qRL[IC];
  const double rRR = qRR[ID], pRR = qRR[IP], uRR = qRR[IU], vRR = qRR[IV], aRR = qRR[IA], bRR = qRR[IB], cRR = qRR[IC];

  double cFastLLx = fast_speed(p, qLL, 0), cFastLRx = fast_speed(p, qLR, 0);
  double cFastRLx = fast_speed(p, qRL, 0), cFastRRx = fast_speed(p, qRR, 0);
  double cFastLLy = fast_speed(p, qLL, 1), cFastLRy = fast_speed(p, qLR, 1);
  double cFastRLy = fast_speed(p, qRL, 1), cFastRRy = fast_speed(p, qRR, 1);

  double SL = min4(uLL, uLR, uRL, uRR) - max4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  double SR = max4(uLL, uLR, uRL, uRR) + max4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  double SB = min4(vLL, vLR, vRL, vRR) - max4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);
  double ST = max4(vLL, vLR, vRL, vRR) + max4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);

  double PtotLL = pLL + 0.5 * (aLL * aLL + bLL * bLL + cLL * cLL);
  double PtotLR = pLR + 0.5 * (aLR * aLR + bLR * bLR + cLR * cLR);
  double PtotRL = pRL + 0.5 * (aRL * aRL + bRL * bRL + cRL * cRL);
  double PtotRR = pRR + 0.5 * (aRR * aRR + bRR * bRR + cRR * cRR);

  double rcLLx = rLL * (uLL - SL), rcRLx = rRL * (SR - uRL), rcLRx = rLR * (uLR - SL), rcRRx = rRR * (SR - uRR);
  double rcLLy = rLL * (vLL - SB), rcLRy = rLR * (ST - vLR), rcRLy = rRL * (vRL - SB), rcRRy = rRR * (ST - vRR);

  double ustar = (rcLLx * uLL + rcLRx * uLR + rcRLx * uRL + rcRRx * uRR + (PtotLL - PtotRL + PtotLR - PtotRR)) /
                 (rcLLx + rcLRx + rcRLx + rcRRx);
  double vstar = (rcLLy * vLL + rcLRy * vLR + rcRLy * vRL + rcRRy * vRR + (PtotLL - PtotLR + PtotRL - PtotRR)) /
                 (rcLLy + rcLRy + rcRLy + rcRRy);

  double rstarLLx = rLL * (SL - uLL) / (SL - ustar);
  double BstarLL = bLL * (SL - uLL) / (SL - ustar);
  double rstarLLy = rLL * (SB - vLL) / (SB - vstar);
  double AstarLL = aLL * (SB - vLL) / (SB - vstar);
  double rstarLL = rLL * (SL - uLL) / (SL - ustar) * (SB - vLL) / (SB - vstar);
  double EstarLLx = ustar * BstarLL - vLL * aLL;
  double EstarLLy = uLL * bLL - vstar * AstarLL;
  double EstarLL = ustar * BstarLL - vstar * AstarLL;

  double rstarLRx = rLR * (SL - uLR) / (SL - ustar);
  double BstarLR = bLR * (SL - uLR) / (SL - ustar);
  double rstarLRy = rLR * (ST - vLR) / (ST - vstar);
  double AstarLR = aLR * (ST - vLR) / (ST - vstar);
  double rstarLR = rLR * (SL - uLR) / (SL - ustar) * (ST - vLR) / (ST - vstar);
  double EstarLRx = ustar * BstarLR - vLR * aLR;
  double EstarLRy = uLR * bLR - vstar * AstarLR;
  double EstarLR = ustar * BstarLR - vstar * AstarLR;

  double rstarRLx = rRL * (SR - uRL) / (SR - ustar);
  double BstarRL = bRL * (SR - uRL) / (SR - ustar);
  double rstarRLy = rRL * (SB - vRL) / (SB - vstar);
  double AstarRL = aRL * (SB - vRL) / (SB - vstar);
  double rstarRL = rRL * (SR - uRL) / (SR - ustar) * (SB - vRL) / (SB - vstar);
  double EstarRLx = ustar * BstarRL - vRL * aRL;
  double EstarRLy = uRL * bRL - vstar * AstarRL;
  double EstarRL = ustar * BstarRL - vstar * AstarRL;

  double rstarRRx = rRR * (SR - uRR) / (SR - ustar);
  double BstarRR = bRR * (SR - uRR) / (SR - ustar);
  double rstarRRy = rRR * (ST - vRR) / (ST - vstar);
  double AstarRR = aRR * (ST - vRR) / (ST - vstar);
  double rstarRR = rRR * (SR - uRR) / (SR - ustar) * (ST - vRR) / (ST - vstar);
  double EstarRRx = ustar * BstarRR - vRR * aRR;
  double EstarRRy = uRR * bRR - vstar * AstarRR;
  double EstarRR = ustar * BstarRR - vstar * AstarRR;

  double calfvenL = max5(fabs(aLR) / sqrt(rstarLRx), fabs(AstarLR) / sqrt(rstarLR), fabs(aLL) / sqrt(rstarLLx),
                         fabs(AstarLL) / sqrt(rstarLL), p->smallc);
  double calfvenR = max5(fabs(aRR) / sqrt(rstarRRx), fabs(AstarRR) / sqrt(rstarRR), fabs(aRL) / sqrt(rstarRLx),
                         fabs(AstarRL) / sqrt(rstarRL), p->smallc);
  double calfvenB = max5(fabs(bLL) / sqrt(rstarLLy), fabs(BstarLL) / sqrt(rstarLL), fabs(bRL) / sqrt(rstarRLy),
                         fabs(BstarRL) / sqrt(rstarRL), p->smallc);
  double calfvenT = max5(fabs(bLR) / sqrt(rstarLRy), fabs(BstarLR) / sqrt(rstarLR), fabs(bRR) / sqrt(rstarRRy),
                         fabs(BstarRR) / sqrt(rstarRR), p->smallc);

  double SAL = fmin(ustar - calfvenL, 0.0);
  double SAR = fmax(ustar + calfvenR, 0.0);
  double SAB = fmin(vstar - calfvenB, 0.0);
  double SAT = fmax(vstar + calfvenT, 0.0);

  double AstarT = (SAR * AstarRR - SAL * AstarLR) / (SAR - SAL);
  double AstarB = (SAR * AstarRL - SAL * AstarLL) / (SAR - SAL);
  double BstarR = (SAT * BstarRR - SAB * BstarRL) / (SAT - SAB);
  double BstarL = (SAT * BstarLR - SAB * BstarLL) / (SAT - SAB);

  /* :561-596 branch-free blend with 0/1 integer weights */
  double E = 0, tmpE = 0;
  int SB_pos = (int)(1 + copysign(1.0, SB)) / 2, SB_neg = 1 - SB_pos;
  int ST_pos = (int)(1 + copysign(1.0, ST)) / 2, ST_neg = 1 - ST_pos;
  int SL_pos = (int)(1 + copysign(1.0, SL)) / 2, SL_neg = 1 - SL_pos;
  int SR_pos = (int)(1 + copysign(1.0, SR)) / 2, SR_neg = 1 - SR_pos;

  tmpE = (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) / (SAR - SAL) / (SAT - SAB) -
         SAT * SAB / (SAT - SAB) * (AstarT - AstarB) + SAR * SAL / (SAR - SAL) * (BstarR - BstarL);
  E += (SB_neg * ST_pos * SL_neg * SR_pos) * tmpE;

  tmpE = (SAR * EstarLLx - SAL * EstarRLx + SAR * SAL * (bRL - bLL)) / (SAR - SAL);
  tmpE = SL_pos * ELL + SL_neg * SR_neg * ERL + SL_neg * SR_pos * tmpE;
  E += SB_pos * tmpE;

  tmpE = (SAR * EstarLRx - SAL * EstarRRx + SAR * SAL * (bRR - bLR)) / (SAR - SAL);
  tmpE = SL_pos * ELR + SL_neg * SR_neg * ERR + SL_neg * SR_pos * tmpE;
  E += (SB_neg * ST_neg) * tmpE;

  tmpE = (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (aLR - aLL)) / (SAT - SAB);
  E += (SB_neg * ST_pos * SL_pos) * tmpE;

  tmpE = (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (aRR - aRL)) / (SAT - SAB);
  E += (SB_neg * ST_pos * SL_neg * SR_neg) * tmpE;

  return E;
}

static double compute_emf(const orc_params *p, state_t qEdge[4], int emfDir)
{
  /* src/shared/RiemannSolvers_MHD.h:651-874 : emfDir 0=EMFX, 1=EMFY, 2=EMFZ.
   * (first parallel, second parallel, orthogonal) = Z:(u,v,w|A,B,C)  Y:(w,u,v|C,A,B)  X:(v,w,u|B,C,A) */
  static const int map_v[3][3] = { { IV, IW, IU }, { IW, IU, IV }, { IU, IV, IW } };
  static const int map_b[3][3] = { { IB, IC, IA }, { IC, IA, IB }, { IA, IB, IC } };
  const int iu = map_v[emfDir][0], iv = map_v[emfDir][1], iw = map_v[emfDir][2];
  const int ia = map_b[emfDir][0], ib = map_b[emfDir][1], ic = map_b[emfDir][2];
  const double *qRT = qEdge[IRT], *qLT = qEdge[ILT], *qRB = qEdge[IRB], *qLB = qEdge[ILB];
  state_t q[4];
  double *qLL = q[ILL], *qRL = q[IRL], *qLR = q[ILR], *qRR = q[IRR];

  qLL[ID] = qRT[ID]; qRL[ID] = qLT[ID]; qLR[ID] = qRB[ID]; qRR[ID] = qLB[ID];
  qLL[IP] = qRT[IP]; qRL[IP] = qLT[IP]; qLR[IP] = qRB[IP]; qRR[IP] = qLB[IP];
  qLL[IU] = qRT[iu]; qRL[IU] = qLT[iu]; qLR[IU] = qRB[iu]; qRR[IU] = qLB[iu];
  qLL[IV] = qRT[iv]; qRL[IV] = qLT[iv]; qLR[IV] = qRB[iv]; qRR[IV] = qLB[iv];
  qLL[IA] = 0.5 * (qRT[ia] + qLT[ia]);
  qRL[IA] = 0.5 * (qRT[ia] + qLT[ia]);
  qLR[IA] = 0.5 * (qRB[ia] + qLB[ia]);
  qRR[IA] = 0.5 * (qRB[ia] + qLB[ia]);
  qLL[IB] = 0.5 * (qRT[ib] + qRB[ib]);
  qRL[IB] = 0.5 * (qLT[ib] + qLB[ib]);
  qLR[IB] = 0.5 * (qRT[ib] + qRB[ib]);
  qRR[IB] = 0.5 * (qLT[ib] + qLB[ib]);
  qLL[IW] = qRT[iw]; qRL[IW] = qLT[iw]; qLR[IW] = qRB[iw]; qRR[IW] = qLB[iw];
  qLL[IC] = qRT[ic]; qRL[IC] = qLT[ic]; qLR[IC] = qRB[ic]; qRR[IC] = qLB[ic];

  double e[4]; /* :835-838 */
  e[ILL] = qLL[IU] * qLL[IB] - qLL[IV] * qLL[IA];
  e[IRL] = qRL[IU] * qRL[IB] - qRL[IV] * qRL[IA];
  e[ILR] = qLR[IU] * qLR[IB] - qLR[IV] * qLR[IA];
  e[IRR] = qRR[IU] * qRR[IB] - qRR[IV] * qRR[IA];
  return mag_riemann2d_hlld(p, q, e);
}

static void compute_emfs(const orc_params *p, orc_scratch *s)
{
  /* src/muscl/MHDRunFunctors3D.h:2181-2230 : range [gw, size-gw+1); note the RB/LT swap for EMF_y */
  const int gw = p->gw;
  double *Emf = s->a3[S_EMF];
#pragma omp parallel for collapse(2)
  for (int k = gw; k < p->ksize - gw + 1; ++k)
    for (int j = gw; j < p->jsize - gw + 1; ++j)
      for (int i = gw; i < p->isize - gw + 1; ++i) {
        state_t qe[4];
        load_state(p, s->a8[S_RT3], i - 1, j - 1, k, qe[IRT]);
        load_state(p, s->a8[S_RB3], i - 1, j, k, qe[IRB]);
        load_state(p, s->a8[S_LT3], i, j - 1, k, qe[ILT]);
        load_state(p, s->a8[S_LB3], i, j, k, qe[ILB]);
        Emf[AT(p, i, j, k, I_EMFZ)] = compute_emf(p, qe, 2);

        load_state(p, s->a8[S_RT2], i - 1, j, k - 1, qe[IRT]);
        load_state(p, s->a8[S_LT2], i, j, k - 1, qe[IRB]);
        load_state(p, s->a8[S_RB2], i - 1, j, k, qe[ILT]);
        load_state(p, s->a8[S_LB2], i, j, k, qe[ILB]);
        Emf[AT(p, i, j, k, I_EMFY)] = compute_emf(p, qe, 1);

        load_state(p, s->a8[S_RT], i, j - 1, k - 1, qe[IRT]);
        load_state(p, s->a8[S_RB], i, j - 1, k, qe[IRB]);
        load_state(p, s->a8[S_LT], i, j, k - 1, qe[ILT]);
        load_state(p, s->a8[S_LB], i, j, k, qe[ILB]);
        Emf[AT(p, i, j, k, I_EMFX)] = compute_emf(p, qe, 0);
      }
}

/* ------------------------------------------------------------------------------------------ */
/* Conservative update + constrained transport                                                */
/* ------------------------------------------------------------------------------------------ */
static void update_hydro(const orc_params *p, double *U, const orc_scratch *s, double dtdx, double dtdy, double dtdz)
{
  /* src/muscl/MHDRunFunctors3D.h:2471-2538 (interior; fixed order; y/z fluxes un-rotated at use).
   * The reference round-trips all 8 variables through a local state; IA..IC are rewritten unchanged. */
  const int gw = p->gw;
  const double *Fx = s->a8[S_FX], *Fy = s->a8[S_FY], *Fz = s->a8[S_FZ];
#pragma omp parallel for collapse(2)
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) {
        state_t u, f;
        load_state(p, U, i, j, k, u);
        load_state(p, Fx, i, j, k, f);
        u[ID] += f[ID] * dtdx; u[IP] += f[IP] * dtdx; u[IU] += f[IU] * dtdx; u[IV] += f[IV] * dtdx; u[IW] += f[IW] * dtdx;
        load_state(p, Fx, i + 1, j, k, f);
        u[ID] -= f[ID] * dtdx; u[IP] -= f[IP] * dtdx; u[IU] -= f[IU] * dtdx; u[IV] -= f[IV] * dtdx; u[IW] -= f[IW] * dtdx;
        load_state(p, Fy, i, j, k, f);
        u[ID] += f[ID] * dtdy; u[IP] += f[IP] * dtdy; u[IU] += f[IV] * dtdy; u[IV] += f[IU] * dtdy; u[IW] += f[IW] * dtdy;
        load_state(p, Fy, i, j + 1, k, f);
        u[ID] -= f[ID] * dtdy; u[IP] -= f[IP] * dtdy; u[IU] -= f[IV] * dtdy; u[IV] -= f[IU] * dtdy; u[IW] -= f[IW] * dtdy;
        load_state(p, Fz, i, j, k, f);
        u[ID] += f[ID] * dtdz; u[IP] += f[IP] * dtdz; u[IU] += f[IW] * dtdz; u[IV] += f[IV] * dtdz; u[IW] += f[IU] * dtdz;
        load_state(p, Fz, i, j, k + 1, f);
        u[ID] -= f[ID] * dtdz; u[IP] -= f[IP] * dtdz; u[IU] -= f[IW] * dtdz; u[IV] -= f[IV] * dtdz; u[IW] -= f[IU] * dtdz;
        store_state(p, U, i, j, k, u);
      }
}

static void update_emf(const orc_params *p, double *U, const orc_scratch *s, double dtdx, double dtdy, double dtdz)
{
  /* src/muscl/MHDRunFunctors3D.h:2581-2622 */
  const int gw = p->gw;
  const double *Emf = s->a3[S_EMF];
#pragma omp parallel for collapse(2)
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) {
        double a = U[AT(p, i, j, k, IA)], b = U[AT(p, i, j, k, IB)], c = U[AT(p, i, j, k, IC)];
        a += (Emf[AT(p, i, j + 1, k, I_EMFZ)] - Emf[AT(p, i, j, k, I_EMFZ)]) * dtdy;
        b -= (Emf[AT(p, i + 1, j, k, I_EMFZ)] - Emf[AT(p, i, j, k, I_EMFZ)]) * dtdx;
        a -= (Emf[AT(p, i, j, k + 1, I_EMFY)] - Emf[AT(p, i, j, k, I_EMFY)]) * dtdz;
        b += (Emf[AT(p, i, j, k + 1, I_EMFX)] - Emf[AT(p, i, j, k, I_EMFX)]) * dtdz;
        c += (Emf[AT(p, i + 1, j, k, I_EMFY)] - Emf[AT(p, i, j, k, I_EMFY)]) * dtdx;
        c -= (Emf[AT(p, i, j + 1, k, I_EMFX)] - Emf[AT(p, i, j, k, I_EMFX)]) * dtdy;
        U[AT(p, i, j, k, IA)] = a;
        U[AT(p, i, j, k, IB)] = b;
        U[AT(p, i, j, k, IC)] = c;
      }
}

void orc_godunov_v0(const orc_params *p, const double *U_in, const double *Q, double *U_out, orc_scratch *s, double dt)
{
  /* src/muscl/SolverMHDMuscl.cpp:477, 490-517 */
  memcpy(U_out, U_in, sizeof(double) * NV * (size_t)orc_ncells(p));
  const double dtdx = dt / p->dx, dtdy = dt / p->dy, dtdz = dt / p->dz;
  compute_elec_field(p, U_in, Q, s->a3[S_ELEC]);
  compute_mag_slopes(p, U_in, s->a3[S_DA], s->a3[S_DB], s->a3[S_DC]);
  compute_trace(p, U_in, Q, s, dtdx, dtdy, dtdz);
  compute_fluxes(p, s);
  compute_emfs(p, s);
  update_hydro(p, U_out, s, dtdx, dtdy, dtdz);
  update_emf(p, U_out, s, dtdx, dtdy, dtdz);
}

double orc_step(const orc_params *p, double *U_in, double *U_out, double *Q, orc_scratch *s, double t, double t_end)
{
  /* src/muscl/SolverMHDMuscl.cpp:465-517 ; SolverBase.cpp:149-179 */
  orc_make_boundaries(p, U_in);
  orc_convert_to_primitives(p, U_in, Q);
  double dt = orc_compute_dt_local(p, Q);
  if (t + dt > t_end) dt = t_end - t;
  orc_godunov_v0(p, U_in, Q, U_out, s, dt);
  return dt;
}

void orc_diagnostics(const orc_params *p, const double *U, double sums[8], double *max_divb)
{
  const int gw = p->gw;
  double m = 0.0;
  for (int v = 0; v < NV; ++v) sums[v] = 0.0;
  for (int k = gw; k < p->ksize - gw; ++k)
    for (int j = gw; j < p->jsize - gw; ++j)
      for (int i = gw; i < p->isize - gw; ++i) {
        for (int v = 0; v < NV; ++v) sums[v] += U[AT(p, i, j, k, v)];
        double d = (U[AT(p, i + 1, j, k, IA)] - U[AT(p, i, j, k, IA)]) / p->dx +
                   (U[AT(p, i, j + 1, k, IB)] - U[AT(p, i, j, k, IB)]) / p->dy +
                   (U[AT(p, i, j, k + 1, IC)] - U[AT(p, i, j, k, IC)]) / p->dz;
        if (fabs(d) > m) m = fabs(d);
      }
  *max_divb = m;
}
