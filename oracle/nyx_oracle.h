/*
 * nyx_oracle.h — TEST INFRASTRUCTURE (see nyx_oracle.c header).  CPU restatement of the
 * reference algorithm; shares only the descriptor PODs of include/nyxb.h with the product.
 */
#ifndef NYX_ORACLE_H
#define NYX_ORACLE_H
#include "../include/nyxb.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct nyx_oracle_grav nyx_oracle_grav;

int nyx_oracle_tableau(int method, int* order, int* stages, const double** a, const double** b);
double nyx_oracle_dur_to_seconds(int64_t total_ns);
int64_t nyx_oracle_dur_from_seconds(double s);
void nyx_oracle_sincos(double x, double* s, double* c);
void nyx_oracle_rotation(const nyxb_rotation* rot, int64_t t_ns, double R[9], double* wdot);
int nyx_oracle_body_position(const nyxb_body* b, int64_t t_ns, double pos[3]);
nyx_oracle_grav* nyx_oracle_grav_new(const nyxb_gravity_field* g);
void nyx_oracle_grav_free(nyx_oracle_grav* h);
void nyx_oracle_grav_accel(const nyx_oracle_grav* h, int64_t t_ns, const double r_in[3], double* scratch, double acc[3]);
double nyx_oracle_occultation(const double r_eb[3], const double r_ls[3], double light_radius_km, double body_radius_km);
double nyx_oracle_error_estimate(int ctrl, const double err[9], const double cand[9], const double cur[9]);
int nyx_oracle_eom(const nyxb_dynamics* dyn, int64_t epoch_ns, double delta_t_s, const double y[9], const double consts[4], double dy[9]);
int nyx_oracle_propagate_batch(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                               const double* state_soa, const double* consts_soa,
                               const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                               double* out_state_soa, int64_t* out_epoch_ns,
                               nyxb_details* out_details, int32_t* out_status, int n_threads);
int nyx_oracle_propagate_batch_event(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                                     const double* state_soa, const double* consts_soa,
                                     const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                     double* out_state_soa, int64_t* out_epoch_ns,
                                     nyxb_details* out_details, int32_t* out_status, const nyxb_traj_sink* sink,
                                     const nyxb_event* event, int n_threads);
int nyx_oracle_propagate_batch_traj(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                                    const double* state_soa, const double* consts_soa,
                                    const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                    double* out_state_soa, int64_t* out_epoch_ns,
                                    nyxb_details* out_details, int32_t* out_status, const nyxb_traj_sink* sink, int n_threads);
/* ---- STM path (nyx_oracle_od.c): dual-number gradient, 90-vector propagation, PropInstance handle ---- */
typedef struct nyx_oracle_inst nyx_oracle_inst;
nyx_oracle_inst* nyx_oracle_inst_new(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, const double y9[9],
                                     const double consts[4], int64_t epoch_ns);
void nyx_oracle_inst_free(nyx_oracle_inst* in);
int nyx_oracle_inst_for_duration(nyx_oracle_inst* in, int64_t duration_ns);
void nyx_oracle_inst_get(const nyx_oracle_inst* in, double y[90], int64_t* epoch_ns, int64_t* step_ns, int* fixed, nyxb_details* det);
void nyx_oracle_inst_set(nyx_oracle_inst* in, const double y[90], int64_t epoch_ns);
void nyx_oracle_inst_set_step(nyx_oracle_inst* in, int64_t step_ns, int fixed);
int nyx_oracle_dual_eom(const nyxb_dynamics* dyn, int64_t t_ns, const double y[9], const double consts[4], double dx[9], double grad[81]);
int nyx_oracle_propagate_batch_stm(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                                   const double* state_soa, const double* consts_soa,
                                   const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                   const double* stm_in_soa, double* out_state_soa, int64_t* out_epoch_ns,
                                   double* out_stm_soa, nyxb_details* out_details, int32_t* out_status, int n_threads);
int nyx_oracle_body_velocity(const nyxb_body* b, int64_t t_ns, double vel[3]);
int nyx_oracle_num_threads(void);
/* sensitivity probe: multiply every adaptive error norm by `s` (1.0 = untouched restatement) */
void nyx_oracle_set_error_scale(double s);
#ifdef __cplusplus
}
#endif
#endif
