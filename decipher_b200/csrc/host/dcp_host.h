/*
 * dcp_host.h — C ABI of the host side of the drop-in: the reference's own file formats in, the
 * reference's own output files out.  Pure C++17 behind it (no CUDA): it restates what `program main`
 * does around the Monte-Carlo loop (RRMODEL/dyna_main.f90:85-160, 166-206):
 *
 *   dcph_load_project   = project + Tread_dyna + Inputs + mc_setup + modelstruct_setup + read_routingdata
 *                         (dyna_project.f90:82-323, dyna_tread_dyna.f90:46-194, dyna_inputs.f90:54-177,
 *                          dyna_mc_setup.f90:81-181, dyna_modelstruct_setup.f90:30-111,
 *                          dyna_read_routingdata.f90:44-102, DTA/dta_riv_tree_node.f90:469-612)
 *   dcph_genpar         = ranin + genpar for members 1..n (dyna_random.f90:51-64,111-133, dyna_genpar.f90:35-41)
 *   dcph_write_member   = file_open + the init block of init_satzone + write_output
 *                         (dyna_file_open.f90:26-41, dyna_init_satzone.f90:177-191, dyna_write_output.f90:44-100)
 *
 * The descriptor returned by dcph_desc() is exactly the dcp_catchment_desc that
 * dcp_catchment_create (include/decipher_b200.h) takes.
 */
#ifndef DCP_HOST_H
#define DCP_HOST_H

#include <stdint.h>

#include "../../../include/decipher_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dcph_model dcph_model;

const char *dcph_last_error(void);

/* workdir: directory that holds project.dat (the reference reads it from the CWD).
 * id_project / id_cat / id_flow: the three integers of the `-auto` file (1-based). */
int dcph_load_project(const char *workdir, int id_project, int id_cat, int id_flow, dcph_model **out);
/* reads the `-auto <file>` triple (dyna_project.f90:125-130) */
int dcph_read_auto_file(const char *path, int *id_project, int *id_cat, int *id_flow);
void dcph_free(dcph_model *m);

const dcp_catchment_desc *dcph_desc(const dcph_model *m);
int dcph_numsim(const dcph_model *m);
void dcph_seeds(const dcph_model *m, int *seed_1, int *seed_2);
/* [npt x 7] column-major lower / upper bounds, parameter order szm lnto srmax srinit chv td smax */
const double *dcph_bounds_ll(const dcph_model *m);
const double *dcph_bounds_ul(const dcph_model *m);
/* output prefix: trim(project_dir)//trim(project_out)//'_'  (dyna_project.f90:322-323) */
const char *dcph_out_prefix(const dcph_model *m);
const char *dcph_log_path(const dcph_model *m);

/* mcpar [npt x 7 x n_member]; the RANMAR stream is restarted from the seeds at every call */
int dcph_genpar(const dcph_model *m, int64_t n_member, double *mcpar);
/* standalone generator entry (known-answer tests) */
void dcph_ranmar(int ij, int kl, int64_t n_skip, int64_t n, double *out);

/* Writes <prefix>mc_id_%08d.res and .flow for member i_mc (1-based).
 *   mcpar [npt x 7] (after the srinit clip), q [nriv x nstep], totals [3 x nriv], diag [3],
 *   status: DCP_ST_* word (its sol_flag selects the initialisation message).
 * f0_leading_zero: 0 = gfortran's F0.d (".50"), 1 = "0.50". */
int dcph_write_member(const dcph_model *m, int64_t i_mc, const double *mcpar, const double *q,
                      const double *totals, const double *diag, int32_t status, int f0_leading_zero);

/* member selection of `decipher_b200_run -write-members 1:10,250` -> ascending unique 0-based indexes; returns the count or -1 */
int64_t dcph_parse_member_list(const char *list, int64_t numsim, int64_t *out, int64_t cap);

/* Standalone router front end (DTA/route_run_standalone.f90): flow tree (*_flow_conn.txt + the *_flow_point.txt beside it), river
 * data and the ASCII-grid flow series (one row per timestep, one column per node -- or one column less when the first node is an
 * ungauged sea outlet).  dcph_desc() then describes the river network (placeholder HRUs) for dcp_catchment_create +
 * dcp_route_timeseries; dcph_router_inflow() is the series in the ABI layout [nriv x nstep]; dcph_write_routed writes the routed
 * series like write_numeric_list (tab-separated gauge ids, then '(1(F0.p) n-1(1x,F0.p))' records). */
int dcph_load_router(const char *river_data_file, const char *flow_conn_file, const char *timeseries_file, dcph_model **out);
const double *dcph_router_inflow(const dcph_model *m);
int dcph_write_routed(const dcph_model *m, const double *routed, const char *path, int precision, int f0_leading_zero);

#ifdef __cplusplus
}
#endif
#endif
