/*
 * oracle/mhd2d_oracle.c  --  TEST INFRASTRUCTURE ONLY (see the header of mhd3d_oracle.h for the rules).
 *
 * Plain-C, CPU restatement of ppkMHD's 2-D MUSCL-Hancock + constrained-transport MHD step, implementationVersion 0
 * (SolverMHDMuscl<2>::godunov_unsplit_impl, src/muscl/SolverMHDMuscl.cpp:373-417): SURVEY 8(f) rank 2, the next row to
 * widen into. There is NO CUDA path for it yet: this file only pins the algorithm, bit for bit, against the unmodified
 * reference (the .npz fixtures under tests/golden2d, tests/test_oracle2d_vs_golden.py), so that the GPU kernels of the next round have a
 * checker from their first line on.
 *
 * One translation unit with the 3-D oracle: the face Riemann solvers (riemann_mhd), the 2-D magnetic HLLD solver
 * (compute_emf) and the primitive-variable conversion are the same functions in 2-D and 3-D in the reference too
 * (src/shared/RiemannSolvers_MHD.h, mhd_utils.h).
 *
 * Layout: U[var][j][i], index = i + isize*(j + jsize*var), ghost width 3, 8 variables, B on the lower faces.
 */
#include "mhd3d_oracle.c"

#define AT2(p, i, j, v) ((size_t)(i) + (size_t)(p)->isize * ((size_t)(j) + (size_t)(p)->jsize * (size_t)(v)))

void orc2d_params_finalize(orc_params *p)
{
  /* src/shared/HydroParams.cpp:421-441, :400-402 with dimType = TWO_D (kmin = kmax = 0, ksize = 1, :119-131) */
  if (p->mx < 1) p->mx = 1;
  if (p->my < 1) p->my = 1;
  p->mz = 1;
  p->nz = 1;
  p->isize = p->nx + 2 * p->gw;
  p->jsize = p->ny + 2 * p->gw;
  p->ksize = 1;
  p->dx = (p->xmax - p->xmin) / (p->nx * p->mx);
  p->dy = (p->ymax - p->ymin) / (p->ny * p->my);
  p->dz = (p->zmax - p->zmin) / (p->nz * p->mz);
  p->smallp = p->smallc * p->smallc / p->gamma0;
}

static void load2(const orc_params *p, const double *A, int i, int j, state_t q)
{
  for (int v = 0; v < NV; ++v) q[v] = A[AT2(p, i, j, v)];
}
static void store2(const orc_params *p, double *A, int i, int j, const state_t q)
{
  for (int v = 0; v < NV; ++v) A[AT2(p, i, j, v)] = q[v];
}

void orc2d_init_orszag_tang(const orc_params *p, double *U)
{
  /* src/muscl/MHDInitFunctors2D.h:241-385 : all variables but the energy, then the energy from the face fields */
  const int gw = p->gw;
  const double TWOPI = 2 * 3.141592653589793238462643383279502884L; /* TWOPI_F, src/shared/real_type.h:65,75 */
  const double B0 = 1.0 / sqrt(2 * TWOPI);
  const double p0 = p->gamma0 / (2 * TWOPI);
  const double d0 = p->gamma0 * p0;
  const double v0 = 1.0;
  for (int j = 0; j < p->jsize; ++j)
    for (int i = 0; i < p->isize; ++i) {
      double xPos = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
      double yPos = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
      U[AT2(p, i, j, ID)] = d0;
      U[AT2(p, i, j, IU)] = -d0 * v0 * sin(yPos * TWOPI);
      U[AT2(p, i, j, IV)] = d0 * v0 * sin(xPos * TWOPI);
      U[AT2(p, i, j, IW)] = 0.0;
      U[AT2(p, i, j, IA)] = -B0 * sin(yPos * TWOPI);
      U[AT2(p, i, j, IB)] = B0 * sin(2.0 * xPos * TWOPI);
      U[AT2(p, i, j, IC)] = 0.0;
      U[AT2(p, i, j, IP)] = 0.0;
    }
  const double TwoPi = 4.0 * asin(1.0);
  const double p0e = p->gamma0 / (2.0 * TwoPi);
  for (int j = 0; j < p->jsize - 1; ++j)
    for (int i = 0; i < p->isize - 1; ++i) {
      double mu = U[AT2(p, i, j, IU)], mv = U[AT2(p, i, j, IV)], d = U[AT2(p, i, j, ID)];
      double bx = U[AT2(p, i, j, IA)] + U[AT2(p, i + 1, j, IA)];
      double by = U[AT2(p, i, j, IB)] + U[AT2(p, i, j + 1, IB)];
      U[AT2(p, i, j, IP)] = p0e / (p->gamma0 - 1.0) + 0.5 * (mu * mu / d + mv * mv / d + 0.25 * (bx * bx) + 0.25 * (by * by));
    }
}

void orc2d_init_blast(const orc_params *p, double radius, double cx, double cy, double density_in, double density_out,
                      double pressure_in, double pressure_out, double *U)
{
  /* InitBlastFunctor2D_MHD, src/muscl/MHDInitFunctors2D.h:144-236 (uniform field 0.5, 0.5, 0.5 hard-coded) */
  const int gw = p->gw;
  const double radius2 = radius * radius;
  for (int j = 0; j < p->jsize; ++j)
    for (int i = 0; i < p->isize; ++i) {
      double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
      double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
      double d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
      int in = d2 < radius2;
      double a = 0.5, b = 0.5, c = 0.5;
      U[AT2(p, i, j, ID)] = in ? density_in : density_out;
      U[AT2(p, i, j, IU)] = 0.0;
      U[AT2(p, i, j, IV)] = 0.0;
      U[AT2(p, i, j, IW)] = 0.0;
      U[AT2(p, i, j, IA)] = a;
      U[AT2(p, i, j, IB)] = b;
      U[AT2(p, i, j, IC)] = c;
      U[AT2(p, i, j, IP)] = (in ? pressure_in : pressure_out) / (p->gamma0 - 1.0) + 0.5 * (a * a + b * b + c * c);
    }
}

void orc2d_init_rotor(const orc_params *p, double r0, double r1, double u0, double p0, double b0, double *U)
{
  /* InitRotorFunctor2D_MHD, src/muscl/MHDInitFunctors2D.h:588-668 (momenta = rho * f_r * u0 * ..., also inside r0) */
  const int gw = p->gw;
  const double xCenter = (p->xmax + p->xmin) / 2, yCenter = (p->ymax + p->ymin) / 2;
  for (int j = 0; j < p->jsize; ++j)
    for (int i = 0; i < p->isize; ++i) {
      double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
      double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
      double r = sqrt((x - xCenter) * (x - xCenter) + (y - yCenter) * (y - yCenter));
      double f_r = (r1 - r) / (r1 - r0);
      double d, mu, mv;
      if (r <= r0) { d = 10.0; mu = -d * f_r * u0 * (y - yCenter) / r0; mv = d * f_r * u0 * (x - xCenter) / r0; }
      else if (r <= r1) { d = 1 + 9 * f_r; mu = -d * f_r * u0 * (y - yCenter) / r; mv = d * f_r * u0 * (x - xCenter) / r; }
      else { d = 1.0; mu = 0.0; mv = 0.0; }
      U[AT2(p, i, j, ID)] = d;
      U[AT2(p, i, j, IU)] = mu;
      U[AT2(p, i, j, IV)] = mv;
      U[AT2(p, i, j, IW)] = 0.0;
      U[AT2(p, i, j, IA)] = b0;
      U[AT2(p, i, j, IB)] = 0.0;
      U[AT2(p, i, j, IC)] = 0.0;
      U[AT2(p, i, j, IP)] = p0 / (p->gamma0 - 1.0) + 0.5 * (mu * mu + mv * mv + 0.0 * 0.0) / d + 0.5 * (b0 * b0);
    }
}

void orc2d_init_field_loop(const orc_params *p, double radius, double density_in, double amplitude, double vflow, double *U)
{
  /* InitFieldLoopFunctor2D_MHD, src/muscl/MHDInitFunctors2D.h:715-945 : A_z everywhere, then the interior cells (face B by
   * first differences of A_z), then their energy; ghost cells stay zero until the first ghost fill */
  const int gw = p->gw, isz = p->isize, jsz = p->jsize, nx = p->nx, ny = p->ny, nz = p->nz;
  double *Az = (double *)calloc((size_t)isz * jsz, sizeof(double));
  memset(U, 0, sizeof(double) * NV * (size_t)isz * jsz);
#define AZ(i, j) Az[(size_t)(i) + (size_t)isz * (size_t)(j)]
  for (int j = 0; j < jsz; ++j)
    for (int i = 0; i < isz; ++i) {
      double x = p->xmin + p->dx / 2 + (i + nx * p->px - gw) * p->dx;
      double y = p->ymin + p->dy / 2 + (j + ny * p->py - gw) * p->dy;
      double r = sqrt(x * x + y * y);
      AZ(i, j) = r < radius ? amplitude * (radius - r) : 0.0;
    }
  const double cos_theta = 2.0 / sqrt(5.0);
  const double sin_theta = sqrt(1 - cos_theta * cos_theta);
  for (int j = gw; j < jsz - gw; ++j)
    for (int i = gw; i < isz - gw; ++i) {
      double x = p->xmin + p->dx / 2 + (i + nx * p->px - gw) * p->dx;
      double y = p->ymin + p->dy / 2 + (j + ny * p->py - gw) * p->dy;
      double diag = sqrt(1.0 * (nx * nx + ny * ny + nz * nz));
      double r = sqrt(x * x + y * y);
      double d = r < radius ? density_in : 1.0;
      U[AT2(p, i, j, ID)] = d;
      U[AT2(p, i, j, IU)] = d * vflow * cos_theta;
      U[AT2(p, i, j, IV)] = d * vflow * sin_theta;
      U[AT2(p, i, j, IW)] = d * vflow * nz / diag;
      U[AT2(p, i, j, IA)] = (AZ(i, j + 1) - AZ(i, j)) / p->dy;
      U[AT2(p, i, j, IB)] = -(AZ(i + 1, j) - AZ(i, j)) / p->dx;
      U[AT2(p, i, j, IC)] = 0.0;
    }
  for (int j = gw; j < jsz - gw; ++j)
    for (int i = gw; i < isz - gw; ++i) {
      double a = U[AT2(p, i, j, IA)] + U[AT2(p, i + 1, j, IA)], b = U[AT2(p, i, j, IB)] + U[AT2(p, i, j + 1, IB)];
      double mu = U[AT2(p, i, j, IU)], mv = U[AT2(p, i, j, IV)];
      U[AT2(p, i, j, IP)] = 1.0f / (p->gamma0 - 1.0) + 0.5 * (0.25 * (a * a) + 0.25 * (b * b)) + 0.5 * (mu * mu + mv * mv) / U[AT2(p, i, j, ID)];
    }
#undef AZ
  free(Az);
}

void orc2d_init_kelvin_helmholtz(const orc_params *p, double d_in, double d_out, double pressure, double vflow_in,
                                 double vflow_out, int mode, double w0, double delta, int sine_robertson, double *U)
{
  /* InitKelvinHelmholtzFunctor2D_MHD, src/muscl/MHDInitFunctors2D.h:394-583, the two deterministic perturbations; the
   * out-of-plane momentum is never written (stays 0) */
  const int gw = p->gw;
  const double PI = 3.141592653589793238462643383279502884L; /* PI_F */
  const double y1 = 0.25, y2 = 0.75;
  for (int j = 0; j < p->jsize; ++j)
    for (int i = 0; i < p->isize; ++i) {
      double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
      double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
      double d, u, v;
      if (sine_robertson) {
        const double ramp = 1.0 / (1.0 + exp(2 * (y - y1) / delta)) + 1.0 / (1.0 + exp(2 * (y2 - y) / delta));
        d = d_in + ramp * (d_out - d_in);
        u = vflow_in + ramp * (vflow_out - vflow_in);
      } else {
        d = (y >= y1 && y <= y2) ? d_in : d_out;
        u = (y >= y1 && y <= y2) ? vflow_in : vflow_out;
      }
      v = w0 * sin(mode * PI * x);
      const double bx = 0.5, by = 0.0, bz = 0.0;
      U[AT2(p, i, j, ID)] = d;
      U[AT2(p, i, j, IU)] = d * u;
      U[AT2(p, i, j, IV)] = d * v;
      U[AT2(p, i, j, IW)] = 0.0;
      U[AT2(p, i, j, IA)] = bx;
      U[AT2(p, i, j, IB)] = by;
      U[AT2(p, i, j, IC)] = bz;
      U[AT2(p, i, j, IP)] = pressure / (p->gamma0 - 1.0) + 0.5 * d * (u * u + v * v) + 0.5 * (bx * bx + by * by + bz * bz);
    }
}

void orc2d_init_implode(const orc_params *p, const double outer[8], const double inner[8], int shape, double *U)
{
  /* InitImplodeFunctor2D_MHD, src/muscl/MHDInitFunctors2D.h:36-139 ; outer/inner = rho, p, u, v, w, Bx, By, Bz of
   * ImplodeParams (w and Bz are not used in 2-D); velocities go into the momentum slots, as in the reference */
  const int gw = p->gw;
  for (int j = 0; j < p->jsize; ++j)
    for (int i = 0; i < p->isize; ++i) {
      double x = p->xmin + p->dx / 2 + (i + p->nx * p->px - gw) * p->dx;
      double y = p->ymin + p->dy / 2 + (j + p->ny * p->py - gw) * p->dy;
      int tmp;
      if (shape == 1) tmp = x + y > 0.5 && x + y < 2.5;
      else tmp = x + y > (p->xmin + p->xmax) / 2. + p->ymin;
      const double *s = tmp ? outer : inner;
      U[AT2(p, i, j, ID)] = s[0];
      U[AT2(p, i, j, IP)] = s[1] / (p->gamma0 - 1.0) + 0.5 * s[0] * (s[2] * s[2] + s[3] * s[3]) + 0.5 * (s[5] * s[5] + s[6] * s[6]);
      U[AT2(p, i, j, IU)] = s[2];
      U[AT2(p, i, j, IV)] = s[3];
      U[AT2(p, i, j, IW)] = 0.0;
      U[AT2(p, i, j, IA)] = s[5];
      U[AT2(p, i, j, IB)] = s[6];
      U[AT2(p, i, j, IC)] = 0.0;
    }
}

void orc2d_make_boundaries(const orc_params *p, double *U)
{
  /* SolverBase::make_boundaries_serial, 2-D branch (src/shared/SolverBase.cpp:505-525) with
   * MakeBoundariesFunctor2D_MHD<face> (src/shared/BoundariesFunctors.h:535-744): XMIN, XMAX, YMIN, YMAX in this order,
   * every face over the full extent of the other direction. Faces with ORC_BC_COPY are left alone. */
  const int gw = p->gw, nx = p->nx, ny = p->ny;
  for (int face = 0; face < 4; ++face) {
    const int bc = p->bc[face];
    if (bc == ORC_BC_COPY) continue;
    const int dir = face / 2, hi = face % 2;
    const int n = dir == 0 ? nx : ny;
    const int other = dir == 0 ? p->jsize : p->isize;
    for (int t = 0; t < other; ++t)
      for (int g = 0; g < gw; ++g) {
        const int c = hi ? g + n + gw : g;
        int c0;
        if (bc == ORC_BC_DIRICHLET) c0 = hi ? 2 * n + 2 * gw - 1 - c : 2 * gw - 1 - c;
        else if (bc == ORC_BC_NEUMANN) c0 = hi ? n + gw - 1 : gw;
        else c0 = hi ? c - n : n + c;
        for (int v = 0; v < NV; ++v) {
          double sign = 1.0;
          if (bc == ORC_BC_DIRICHLET && (v == IU + dir || v == IA + dir)) sign = -1.0;
          if (dir == 0) U[AT2(p, c, t, v)] = U[AT2(p, c0, t, v)] * sign;
          else U[AT2(p, t, c, v)] = U[AT2(p, t, c0, v)] * sign;
        }
      }
  }
}

void orc2d_convert_to_primitives(const orc_params *p, const double *U, double *Q)
{
  /* src/muscl/MHDRunFunctors2D.h:85-160 : the out-of-plane neighbour field is 0.0 (not the cell's own Bz) */
  for (int j = 0; j < p->jsize - 1; ++j)
    for (int i = 0; i < p->isize - 1; ++i) {
      state_t u, q;
      double bn[3];
      load2(p, U, i, j, u);
      bn[0] = U[AT2(p, i + 1, j, IA)];
      bn[1] = U[AT2(p, i, j + 1, IB)];
      bn[2] = 0.0;
      constoprim(p, u, bn, q);
      store2(p, Q, i, j, q);
    }
}

double orc2d_compute_inv_dt(const orc_params *p, const double *Q)
{
  /* src/muscl/MHDRunFunctors2D.h:16-83 with find_speed_info<TWO_D> (src/shared/mhd_utils.h:319-366) */
  const int gw = p->gw;
  double invDt = 0.0;
  for (int j = gw; j < p->jsize - gw; ++j)
    for (int i = gw; i < p->isize - gw; ++i) {
      state_t q;
      load2(p, Q, i, j, q);
      double vx = fast_speed(p, q, 0) + fabs(q[IU]);
      double vy = fast_speed(p, q, 1) + fabs(q[IV]);
      invDt = fmax(invDt, vx / p->dx + vy / p->dy);
    }
  return invDt;
}

static void floor2d(const orc_params *p, state_t s)
{
  /* 2-D floors: pressure against smallp * rho (MHDBaseFunctor2D.h:1004-1005), unlike the 3-D v0 trace */
  s[ID] = fmax(p->smallr, s[ID]);
  s[IP] = fmax(p->smallp * s[ID], s[IP]);
}

static void mag_slopes2d(const orc_params *p, const double bf[6], double *dbfY_x, double *dbfX_y)
{
  /* slope_unsplit_mhd_2d, src/muscl/MHDBaseFunctor2D.h:456-520 : bf = bfx, bfx(y+), bfx(y-), bfy, bfy(x+), bfy(x-) */
  *dbfY_x = limited_slope(p->slope_type, bf[0], bf[1], bf[2]);
  *dbfX_y = limited_slope(p->slope_type, bf[3], bf[4], bf[5]);
}

/* trace_unsplit_mhd_2d, src/muscl/MHDBaseFunctor2D.h:776-1109. qNb[di][dj] = Q(i+di-1, j+dj-1),
 * bfx/bfy[di][dj] = face fields U_A / U_B at (i+di-1, j+dj-1), di, dj in 0..3. */
static void trace_cell2d(const orc_params *p, state_t qNb[3][3], double bfx[4][4], double bfy[4][4], double dtdx,
                         double dtdy, state_t qm[2], state_t qp[2], state_t qEdge[4])
{
  enum { C = 1 };
  const double gamma = p->gamma0;
  const double *q = qNb[C][C];
  double Ez[2][2];
  for (int di = 0; di < 2; ++di)
    for (int dj = 0; dj < 2; ++dj) {
      int cx = C + di, cy = C + dj;
      double u = 0.25 * (qNb[cx - 1][cy - 1][IU] + qNb[cx - 1][cy][IU] + qNb[cx][cy - 1][IU] + qNb[cx][cy][IU]);
      double v = 0.25 * (qNb[cx - 1][cy - 1][IV] + qNb[cx - 1][cy][IV] + qNb[cx][cy - 1][IV] + qNb[cx][cy][IV]);
      double A = 0.5 * (bfx[cx][cy - 1] + bfx[cx][cy]);
      double B = 0.5 * (bfy[cx - 1][cy] + bfy[cx][cy]);
      Ez[di][dj] = u * B - v * A;
    }
  const double ELL = Ez[0][0], ELR = Ez[0][1], ERL = Ez[1][0], ERR = Ez[1][1];
  double r = q[ID], pr = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], Cc = q[IC];
  double AL = bfx[C][C], AR = bfx[C + 1][C], BL = bfy[C][C], BR = bfy[C][C + 1];

  /* hydro slopes (slope_unsplit_hydro_2d :347-404), halved (:870-903) */
  double dqx[NV], dqy[NV];
  const int lim = (p->slope_type == 1 || p->slope_type == 2);
  for (int n = 0; n < NV; ++n) {
    dqx[n] = lim ? limited_slope(p->slope_type, q[n], qNb[C + 1][C][n], qNb[C - 1][C][n]) : 0.0;
    dqy[n] = lim ? limited_slope(p->slope_type, q[n], qNb[C][C + 1][n], qNb[C][C - 1][n]) : 0.0;
    dqx[n] *= 0.5;
    dqy[n] *= 0.5;
  }
  const double drx = dqx[ID], dpx = dqx[IP], dux = dqx[IU], dvx = dqx[IV], dwx = dqx[IW], dCx = dqx[IC], dBx = dqx[IB];
  const double dry = dqy[ID], dpy = dqy[IP], duy = dqy[IU], dvy = dqy[IV], dwy = dqy[IW], dCy = dqy[IC], dAy = dqy[IA];

  /* face-centred transverse slopes at (i,j), (i+1,j), (i,j+1) (:908-953) */
  double bf[6], sy_x, sx_y;
  bf[0] = bfx[C][C]; bf[1] = bfx[C][C + 1]; bf[2] = bfx[C][C - 1];
  bf[3] = bfy[C][C]; bf[4] = bfy[C + 1][C]; bf[5] = bfy[C - 1][C];
  mag_slopes2d(p, bf, &sy_x, &sx_y);
  const double dALy = 0.5 * sy_x, dBLx = 0.5 * sx_y;
  bf[0] = bfx[C + 1][C]; bf[1] = bfx[C + 1][C + 1]; bf[2] = bfx[C + 1][C - 1];
  bf[3] = bfy[C + 1][C]; bf[4] = bfy[C + 2][C]; bf[5] = bfy[C][C];
  mag_slopes2d(p, bf, &sy_x, &sx_y);
  const double dARy = 0.5 * sy_x;
  bf[0] = bfx[C][C + 1]; bf[1] = bfx[C][C + 2]; bf[2] = bfx[C][C];
  bf[3] = bfy[C][C + 1]; bf[4] = bfy[C + 1][C + 1]; bf[5] = bfy[C - 1][C + 1];
  mag_slopes2d(p, bf, &sy_x, &sx_y);
  const double dBRx = 0.5 * sx_y;

  const double dAx = 0.5 * (AR - AL), dBy = 0.5 * (BR - BL);

  /* source terms (:963-991) */
  const double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy;
  const double su0 = (-u * dux - dpx / r - B * dBx / r - Cc * dCx / r) * dtdx + (-v * duy + B * dAy / r) * dtdy;
  const double sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - dpy / r - A * dAy / r - Cc * dCy / r) * dtdy;
  const double sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy;
  const double sp0 = (-u * dpx - dux * gamma * pr) * dtdx + (-v * dpy - dvy * gamma * pr) * dtdy;
  const double sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
  const double sB0 = (-u * dBx - B * dux + v * dAx + A * dvx) * dtdx;
  const double sC0 = (w * dAx + A * dwx - u * dCx - Cc * dux) * dtdx + (-v * dCy - Cc * dvy + w * dBy + B * dwy) * dtdy;
  const double sAL0 = +(ELR - ELL) * 0.5 * dtdy;
  const double sAR0 = +(ERR - ERL) * 0.5 * dtdy;
  const double sBL0 = -(ERL - ELL) * 0.5 * dtdx;
  const double sBR0 = -(ERR - ELR) * 0.5 * dtdx;

  r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; pr = pr + sp0; A = A + sA0; B = B + sB0; Cc = Cc + sC0;
  AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0;

#define SET2(S, R_, U_, V_, W_, P_, A_, B_, C_) \
  do { (S)[ID] = (R_); (S)[IU] = (U_); (S)[IV] = (V_); (S)[IW] = (W_); (S)[IP] = (P_); (S)[IA] = (A_); (S)[IB] = (B_); \
       (S)[IC] = (C_); floor2d(p, (S)); } while (0)
  SET2(qp[0], r - drx, u - dux, v - dvx, w - dwx, pr - dpx, AL, B - dBx, Cc - dCx);       /* right state at left interface */
  SET2(qm[0], r + drx, u + dux, v + dvx, w + dwx, pr + dpx, AR, B + dBx, Cc + dCx);       /* left state at right interface */
  SET2(qp[1], r - dry, u - duy, v - dvy, w - dwy, pr - dpy, A - dAy, BL, Cc - dCy);       /* top state at bottom interface */
  SET2(qm[1], r + dry, u + duy, v + dvy, w + dwy, pr + dpy, A + dAy, BR, Cc + dCy);       /* bottom state at top interface */
  SET2(qEdge[IRT], r + (+drx + dry), u + (+dux + duy), v + (+dvx + dvy), w + (+dwx + dwy), pr + (+dpx + dpy),
       AR + (+dARy), BR + (+dBRx), Cc + (+dCx + dCy));
  SET2(qEdge[IRB], r + (+drx - dry), u + (+dux - duy), v + (+dvx - dvy), w + (+dwx - dwy), pr + (+dpx - dpy),
       AR + (-dARy), BL + (+dBLx), Cc + (+dCx - dCy));
  SET2(qEdge[ILB], r + (-drx - dry), u + (-dux - duy), v + (-dvx - dvy), w + (-dwx - dwy), pr + (-dpx - dpy),
       AL + (-dALy), BL + (-dBLx), Cc + (-dCx - dCy));
  SET2(qEdge[ILT], r + (-drx + dry), u + (-dux + duy), v + (-dvx + dvy), w + (-dwx + dwy), pr + (-dpx + dpy),
       AL + (+dALy), BR + (-dBRx), Cc + (-dCx + dCy));
#undef SET2
}

/* godunov_unsplit_impl<2>, v0 branch, AFTER make_boundaries, convertToPrimitives and compute_dt
 * (src/muscl/SolverMHDMuscl.cpp:383-417): U_out = U_in; trace; fluxes; emf; update; CT update */
void orc2d_godunov_v0(const orc_params *p, const double *U_in, const double *Q, double *U_out, double dt)
{
  const int gw = p->gw, isz = p->isize, jsz = p->jsize;
  const size_t n = (size_t)isz * jsz;
  const double dtdx = dt / p->dx, dtdy = dt / p->dy;
  enum { A_QMX, A_QMY, A_QPX, A_QPY, A_RT, A_RB, A_LT, A_LB, A_FX, A_FY, A_N };
  double *a8[A_N];
  for (int a = 0; a < A_N; ++a) a8[a] = (double *)calloc(NV * n, sizeof(double));
  double *emf = (double *)calloc(n, sizeof(double));
  memcpy(U_out, U_in, sizeof(double) * NV * n);

  /* ComputeTraceFunctor2D_MHD, src/muscl/MHDRunFunctors2D.h:782-915 */
  for (int j = gw - 2; j < jsz - gw + 1; ++j)
    for (int i = gw - 2; i < isz - gw + 1; ++i) {
      state_t qNb[3][3], qm[2], qp[2], qEdge[4];
      double bfx[4][4], bfy[4][4];
      for (int di = 0; di < 3; ++di)
        for (int dj = 0; dj < 3; ++dj) load2(p, Q, i + di - 1, j + dj - 1, qNb[di][dj]);
      for (int di = 0; di < 4; ++di)
        for (int dj = 0; dj < 4; ++dj) {
          bfx[di][dj] = U_in[AT2(p, i + di - 1, j + dj - 1, IA)];
          bfy[di][dj] = U_in[AT2(p, i + di - 1, j + dj - 1, IB)];
        }
      trace_cell2d(p, qNb, bfx, bfy, dtdx, dtdy, qm, qp, qEdge);
      store2(p, a8[A_QMX], i, j, qm[0]); store2(p, a8[A_QPX], i, j, qp[0]);
      store2(p, a8[A_QMY], i, j, qm[1]); store2(p, a8[A_QPY], i, j, qp[1]);
      store2(p, a8[A_RT], i, j, qEdge[IRT]); store2(p, a8[A_RB], i, j, qEdge[IRB]);
      store2(p, a8[A_LT], i, j, qEdge[ILT]); store2(p, a8[A_LB], i, j, qEdge[ILB]);
    }
  /* ComputeFluxesAndStoreFunctor2D_MHD (:375-470) and ComputeEmfAndStoreFunctor2D (:612-686) */
  for (int j = gw; j < jsz - gw + 1; ++j)
    for (int i = gw; i < isz - gw + 1; ++i) {
      state_t ql, qr, f;
      load2(p, a8[A_QMX], i - 1, j, ql);
      load2(p, a8[A_QPX], i, j, qr);
      riemann_mhd(p, ql, qr, f);
      store2(p, a8[A_FX], i, j, f);
      load2(p, a8[A_QMY], i, j - 1, ql);
      swap2(&ql[IU], &ql[IV]); swap2(&ql[IA], &ql[IB]);
      load2(p, a8[A_QPY], i, j, qr);
      swap2(&qr[IU], &qr[IV]); swap2(&qr[IA], &qr[IB]);
      riemann_mhd(p, ql, qr, f);
      store2(p, a8[A_FY], i, j, f);
      state_t qe[4];
      load2(p, a8[A_RT], i - 1, j - 1, qe[IRT]);
      load2(p, a8[A_RB], i - 1, j, qe[IRB]);
      load2(p, a8[A_LT], i, j - 1, qe[ILT]);
      load2(p, a8[A_LB], i, j, qe[ILB]);
      emf[(size_t)i + (size_t)isz * j] = compute_emf(p, qe, 2);
    }
  /* UpdateFunctor2D_MHD (:1539-1640: rho, E, momenta and Bz) then UpdateEmfFunctor2D (:1646-1692) */
  for (int j = gw; j < jsz - gw; ++j)
    for (int i = gw; i < isz - gw; ++i) {
      state_t u, f;
      load2(p, U_out, i, j, u);
      load2(p, a8[A_FX], i, j, f);
      u[ID] += f[ID] * dtdx; u[IP] += f[IP] * dtdx; u[IU] += f[IU] * dtdx; u[IV] += f[IV] * dtdx; u[IW] += f[IW] * dtdx; u[IC] += f[IC] * dtdx;
      load2(p, a8[A_FX], i + 1, j, f);
      u[ID] -= f[ID] * dtdx; u[IP] -= f[IP] * dtdx; u[IU] -= f[IU] * dtdx; u[IV] -= f[IV] * dtdx; u[IW] -= f[IW] * dtdx; u[IC] -= f[IC] * dtdx;
      load2(p, a8[A_FY], i, j, f);
      u[ID] += f[ID] * dtdy; u[IP] += f[IP] * dtdy; u[IU] += f[IV] * dtdy; u[IV] += f[IU] * dtdy; u[IW] += f[IW] * dtdy; u[IC] += f[IC] * dtdy;
      load2(p, a8[A_FY], i, j + 1, f);
      u[ID] -= f[ID] * dtdy; u[IP] -= f[IP] * dtdy; u[IU] -= f[IV] * dtdy; u[IV] -= f[IU] * dtdy; u[IW] -= f[IW] * dtdy; u[IC] -= f[IC] * dtdy;
      store2(p, U_out, i, j, u);
    }
  for (int j = gw; j < jsz - gw; ++j)
    for (int i = gw; i < isz - gw; ++i) {
      const double e = emf[(size_t)i + (size_t)isz * j];
      U_out[AT2(p, i, j, IA)] += (emf[(size_t)i + (size_t)isz * (j + 1)] - e) * dtdy;
      U_out[AT2(p, i, j, IB)] -= (emf[(size_t)(i + 1) + (size_t)isz * j] - e) * dtdx;
    }
  for (int a = 0; a < A_N; ++a) free(a8[a]);
  free(emf);
}

double orc2d_step(const orc_params *p, double *U_in, double *U_out, double *Q, double t, double t_end)
{
  /* src/muscl/SolverMHDMuscl.cpp:373-417 ; SolverBase.cpp:149-179 */
  orc2d_make_boundaries(p, U_in);
  orc2d_convert_to_primitives(p, U_in, Q);
  double dt = p->cfl / orc2d_compute_inv_dt(p, Q);
  if (t + dt > t_end) dt = t_end - t;
  orc2d_godunov_v0(p, U_in, Q, U_out, dt);
  return dt;
}
