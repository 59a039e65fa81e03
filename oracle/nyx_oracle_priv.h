/* nyx_oracle_priv.h — TEST INFRASTRUCTURE: internals shared by nyx_oracle.c and nyx_oracle_od.c. */
#ifndef NYX_ORACLE_PRIV_H
#define NYX_ORACLE_PRIV_H
#include "nyx_oracle.h"
/* GravityField precomputed tables — dynamics/gravity_field.rs:36-48, 52-132 */
struct nyx_oracle_grav {
    int n, m;                 /* max degree / order */
    int dim;                  /* n + 3 */
    double mu, r_eq;
    nyxb_rotation rot;
    double *a_diag;           /* gravity_field.rs:61-66: a_nm[(k,k)], k = 0..n+2 */
    double *b_nm, *c_nm, *vr01, *vr11; /* (n+2)^2 each, row-major [n][m], gravity_field.rs:69-92 */
    double *cbar, *sbar;      /* (n+1)^2 row-major */
};

extern double nyx_oracle_error_scale_;   /* sensitivity probe: factor on every adaptive error norm (1.0 = untouched) */
#endif
