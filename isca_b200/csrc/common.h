// common.h -- shared host/device declarations of the isca_b200 library.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/isca_b200.h"

#define ISCA_KMAX 80          // max num_levels supported by the column kernels (registers/local arrays)

namespace isca {

// ---------------------------------------------------------------------------------------------
// Geometry: sizes, decomposition and the packed triangular spectral layout.
//
// Spectral fields live on the device in a PACKED layout: for every zonal wavenumber m owned by
// this rank, rows n = 0 .. Nm-1 with Nm = M - m + 2 (the triangular truncation m+n <= M plus the
// one extra row the degree-raising operators need, spherical.F90:38-42 / SURVEY app. A), each
// row holding `L` complex levels contiguously:   spec[(off[mi] + n) * L + lev]  (double2).
// Everything outside this triangle is identically zero in the reference's rectangular arrays.
// ---------------------------------------------------------------------------------------------
struct Geometry {
  int I, J, K, M, N;          // lon_max, lat_max, num_levels, num_fourier, num_spherical (= M+1)
  int Jh;                     // J / 2
  int P, rank;                // number of ranks, this rank
  int Jloc, j0;               // local latitude block [j0, j0+Jloc)
  int nm;                     // number of zonal wavenumbers owned by this rank
  int T;                      // packed (m,n) rows owned by this rank
  int nm_max;                 // max nm over ranks
  std::vector<int> m_of;      // [nm]   global m of local index mi
  std::vector<int> off;       // [nm+1] packed row offset of local mi
  std::vector<int> owner;     // [M+1]  rank owning m
  std::vector<int> pos;       // [M+1]  position of m in the lat-owner-side Fourier layout
  std::vector<int> nm_rank;   // [P]    nm of every rank
  std::vector<int> roff;      // [P+1]  prefix sum of nm_rank
};

// device-side copy of the small integer tables
struct GeomDev {
  int I, J, K, M, N, Jh, P, rank, Jloc, j0, nm, T;
  const int* m_of;            // [nm]
  const int* off;             // [nm+1]
  const int* pos;             // [M+1]
  const int* row_m;           // [T] local mi of packed row
  // peer-memory transpose (one process per GPU, buffers mapped with CUDA IPC): when p2p != 0 the Legendre
  // epilogue stores straight into the destination rank's lat-owner buffer and the FFT epilogue straight into the
  // destination rank's m-owner buffer, over NVLink; no send/recv, no staging copy.
  int p2p;
  int symmetric;              // make_symmetric: the truncation also removes every m > 0
  const int* owner;           // [M+1] rank owning m
  const int* lidx;            // [M+1] local index of m on its owner
  const int* nm_rank;         // [P]
  const int* roff;            // [P+1]
  double* const* peerA;       // [P] m-owner-layout Fourier buffer of every rank
  double* const* peerB;       // [P] lat-owner-layout Fourier buffer of every rank
};

// host-computed double tables (see host_tables.cpp)
struct HostTables {
  std::vector<double> sin_hem, wts_hem;              // [Jh]
  std::vector<double> sin_lat, cos_lat, cosm_lat, cosm2_lat, wts_lat, deg_lat, rad_lat, coriolis; // [J]
  std::vector<double> deg_lon;                       // [I]
  std::vector<double> pk, bk, dpk, dbk;              // [K+1],[K+1],[K],[K]
  std::vector<double> leg, legw;                     // packed [T][Jh]: P(m,n,jh), P*w
  // per packed row coefficient tables [T]
  std::vector<double> eigen, coef_uvm, coef_uvc, coef_uvp, coef_alpm, coef_alpp, coef_dym, coef_dx, coef_dyp;
  std::vector<double> trunc_mask;                    // 1 where m+n <= M else 0
  std::vector<double> damping, damping_vor, damping_div, eddy_sponge, zmu_sponge, zmv_sponge; // [T]
  std::vector<int>    row_m, row_n;                  // [T]
  // semi-implicit reference state (implicit.F90)
  std::vector<double> ref_ln_p_half, ref_ln_p_full, ref_t, h, div_mat;  // [K+1],[K],[K],[K],[K*K]
  // finite-volume tracer grid (fv_advection_init): see tracer.h
  std::vector<double> fv_c, fv_cc, fv_dy, fv_dyy, fv_dy_plus, fv_dy_minus; double fv_dx;
  std::vector<double> twiddle;                       // [I/2][2]: exp(-2 pi i k / I), k < I/2
  double ref_ps;
  double global_sum_of_wts;
};

void build_geometry(const IscaConfig& c, int rank, int nranks, Geometry& g);
void build_tables(const IscaConfig& c, const Geometry& g, HostTables& t);
// wave_matrix(:,:,L) = (I + xi^2 L(L+1)/a^2 div_mat)^-1 for L = 0..M (implicit.F90:218-237), row-major [L][k][kk]
void build_wave_matrices(const IscaConfig& c, const Geometry& g, const HostTables& t, double xi,
                         std::vector<double>& wm);
void compute_gaussian(int n_hem, std::vector<double>& sin_hem, std::vector<double>& wts_hem);
void invert_matrix(std::vector<double>& a, int n);   // matrix_invert.F90:38-130, row-major in/out

}  // namespace isca
