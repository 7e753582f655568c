// decipher_main.cpp — command-line driver with the reference executable's interface:
//
//     cd <dir with project.dat> && decipher_b200_run -auto <file>
//
// mirrors `program main` (RRMODEL/dyna_main.f90:85-206): reads project.dat / settings / params /
// model_structure and the DTA files, samples every parameter set with RANMAR (bit-exact with the
// reference, which draws them inside the loop from the same single stream), then makes ONE
// dcp_run_ensemble call on the GPU instead of the serial `do i_mc` loop and writes the same
// mc_id_%08d.res / .flow files plus a binary table of performance measures.
#include "dcp_host.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void usage() {
    fprintf(stderr,
            "usage: decipher_b200_run [-auto <file>] [-ids P C F] [-dir <workdir>] [-members N] [-no-flow] [-fast] [-check]\n"
            "  -auto <file>   three lines: project id, catchment-setup id, input-setup id (as the reference)\n"
            "  -ids P C F     the same three ids on the command line\n"
            "  -dir <workdir> directory holding project.dat (default: current directory)\n"
            "  -members N     override numsim of params.dat\n"
            "  -no-flow       do not write per-member .res/.flow text files (large ensembles)\n"
            "  -fast          DCP_OPT_FAST_MATH;  -check  DCP_OPT_CHECK_BALANCE (results_ts STOP conditions as flags)\n");
}

int main(int argc, char **argv) {
    printf(" --- Starting DECIPHeR rainfall-runoff modelling (B200 ensemble path) ---\n");
    std::string workdir = ".", auto_file;
    int ids[3] = {0, 0, 0};
    long members = -1;
    bool write_text = true;
    int flags = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-auto" && i + 1 < argc) auto_file = argv[++i];
        else if (a == "-ids" && i + 3 < argc) { ids[0] = atoi(argv[++i]); ids[1] = atoi(argv[++i]); ids[2] = atoi(argv[++i]); }
        else if (a == "-dir" && i + 1 < argc) workdir = argv[++i];
        else if (a == "-members" && i + 1 < argc) members = atol(argv[++i]);
        else if (a == "-no-flow") write_text = false;
        else if (a == "-fast") flags |= DCP_OPT_FAST_MATH;
        else if (a == "-check") flags |= DCP_OPT_CHECK_BALANCE;
        else { usage(); return 2; }
    }
    if (!auto_file.empty()) {
        if (dcph_read_auto_file(auto_file.c_str(), &ids[0], &ids[1], &ids[2]) != DCP_OK) {
            fprintf(stderr, "cannot read -auto file: %s\n", dcph_last_error());
            return 1;
        }
    }
    if (ids[0] < 1) {   // the reference prompts interactively (dyna_project.f90:133,198,233)
        printf(" Input the project, HRU file and INPUT file ids you wish to use for this run?\n");
        if (scanf("%d %d %d", &ids[0], &ids[1], &ids[2]) != 3) { usage(); return 2; }
    }
    dcph_model *M = nullptr;
    if (dcph_load_project(workdir.c_str(), ids[0], ids[1], ids[2], &M) != DCP_OK) {
        fprintf(stderr, "setup failed: %s\n", dcph_last_error());
        return 1;
    }
    const dcp_catchment_desc *d = dcph_desc(M);
    const long numsim = members > 0 ? members : dcph_numsim(M);
    printf(" %d HRUs, %d rivers, %d timesteps, %d parameter layers, %ld simulations\n", d->nac, d->nriv, d->nstep, d->npt, numsim);
    FILE *log = fopen(dcph_log_path(M), "w");
    if (log) fprintf(log, " --- DYMOND ---\n \n Reading in project files\n Number of HRUs %d\n", d->nac);

    if (dcp_device_count() < 1) { fprintf(stderr, "no CUDA device: this build has no CPU fallback\n"); return 1; }
    dcp_catchment *cat = nullptr;
    if (dcp_catchment_create(d, &cat) != DCP_OK) { fprintf(stderr, "dcp_catchment_create: %s\n", dcp_last_error()); return 1; }

    const size_t npar = (size_t)d->npt * 7;
    std::vector<double> mcpar(npar * numsim), perf((size_t)DCP_NMEASURE * d->nriv * numsim),
        totals((size_t)3 * d->nriv * numsim), diag((size_t)3 * numsim);
    std::vector<int32_t> status(numsim);
    dcph_genpar(M, numsim, mcpar.data());
    if (log) fprintf(log, " \n Looping through simulations...\n");

    dcp_run_opts opts{};
    opts.struct_size = sizeof(opts);
    opts.flags = flags;
    const auto t0 = std::chrono::steady_clock::now();
    // text output needs the flow series on the host: run those members in chunks that fit in memory
    const long chunk = write_text ? std::max<long>(1, std::min<long>(numsim, (long)(2e9 / (8.0 * d->nriv * d->nstep)))) : numsim;
    std::vector<double> q;
    if (write_text) q.resize((size_t)d->nriv * d->nstep * chunk);
    for (long m0 = 0; m0 < numsim; m0 += chunk) {
        const long n = std::min(chunk, numsim - m0);
        const int rc = dcp_run_ensemble(cat, m0 + 1, n, mcpar.data() + npar * m0, &opts, write_text ? q.data() : nullptr,
                                        perf.data() + (size_t)DCP_NMEASURE * d->nriv * m0, totals.data() + (size_t)3 * d->nriv * m0,
                                        diag.data() + (size_t)3 * m0, status.data() + m0);
        if (rc != DCP_OK) { fprintf(stderr, "dcp_run_ensemble: %s\n", dcp_last_error()); return 1; }
        if (write_text)
            for (long m = 0; m < n; ++m) {
                printf(" Running sim %12ld\n", m0 + m + 1);
                if (dcph_write_member(M, m0 + m + 1, mcpar.data() + npar * (m0 + m), q.data() + (size_t)d->nriv * d->nstep * m,
                                      totals.data() + (size_t)3 * d->nriv * (m0 + m), diag.data() + (size_t)3 * (m0 + m),
                                      status[m0 + m], 0) != DCP_OK) {
                    fprintf(stderr, "write_output: %s\n", dcph_last_error());
                    return 1;
                }
            }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // binary performance table: header (int64 numsim, nriv, nmeasure) then perf, status
    const std::string table = std::string(dcph_out_prefix(M)) + "ensemble_perf.bin";
    if (FILE *f = fopen(table.c_str(), "wb")) {
        const int64_t hdr[3] = {numsim, d->nriv, DCP_NMEASURE};
        fwrite(hdr, 8, 3, f);
        fwrite(perf.data(), 8, perf.size(), f);
        fwrite(status.data(), 4, status.size(), f);
        fclose(f);
    }
    long n_flag = 0;
    for (long m = 0; m < numsim; ++m) if (status[m] & ~DCP_ST_SOLFLAG_MASK) ++n_flag;
    printf(" %ld simulations, %.3e member-HRU-timesteps/s (%.2f s); %ld members carry a status flag\n", numsim,
           (double)numsim * d->nac * d->nstep / secs, secs, n_flag);
    if (log) { fprintf(log, " done: %ld simulations in %.2f s\n", numsim, secs); fclose(log); }
    dcp_catchment_destroy(cat);
    dcph_free(M);
    return 0;
}
