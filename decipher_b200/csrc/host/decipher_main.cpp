// decipher_main.cpp — command-line driver with the reference executable's interface:
//
//     cd <dir with project.dat> && decipher_b200_run -auto <file>
//
// mirrors `program main` (RRMODEL/dyna_main.f90:85-206): reads project.dat / settings / params /
// model_structure and the DTA files, samples every parameter set with RANMAR (bit-exact with the
// reference, which draws them inside the loop from the same single stream), then makes ONE
// dcp_run_ensemble call on the GPU (dcp_run_ensemble_multi over -gpus N devices) instead of the serial `do i_mc`
// loop and writes the same mc_id_%08d.res / .flow files plus a binary table of parameters and performance measures.
//
// Output at ensemble scale (the reference writes two text files per member,
// RRMODEL/dyna_write_output.f90:44-100, dyna_file_open.f90:26-41):
//   -no-flow               measures only: <prefix>ensemble_perf.bin, nothing per member
//   -write-members a:b,c   text .res/.flow files for the selected members only (1-based, ranges inclusive)
//   -flow-bin              the selected members' flow series as FP64 in <prefix>ensemble_flow.bin (text .flow keeps
//                          ~4 significant digits of m/timestep flows; the binary file carries every bit)
//   ensemble_perf.bin v2   header {magic "DCPPERF2", int64 numsim, nriv, nmeasure, npt}, then mcpar[npt x 7 x numsim]
//                          (AFTER the srinit clip, the values the .res table prints), perf[nmeasure x nriv x numsim],
//                          status[numsim] -- the sampled parameter sets sit next to their measures
#include "dcp_host.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void usage() {
    fprintf(stderr,
            "usage: decipher_b200_run [-auto <file>] [-ids P C F] [-dir <workdir>] [-members N] [-no-flow] [-write-members LIST]\n"
            "                         [-flow-bin] [-gpus N] [-fast] [-check]\n"
            "  -auto <file>   three lines: project id, catchment-setup id, input-setup id (as the reference)\n"
            "  -ids P C F     the same three ids on the command line\n"
            "  -dir <workdir> directory holding project.dat (default: current directory)\n"
            "  -members N     override numsim of params.dat\n"
            "  -no-flow       do not write per-member .res/.flow text files (large ensembles)\n"
            "  -write-members LIST  per-member files for these members only, e.g. 1:10,250,9000:9003 (1-based, inclusive)\n"
            "  -flow-bin      also write the selected members' flows as FP64 to <prefix>ensemble_flow.bin\n"
            "  -gpus N        spread the members over N devices (dcp_run_ensemble_multi; one gather of the measures)\n"
            "  -fast          DCP_OPT_FAST_MATH;  -check  DCP_OPT_CHECK_BALANCE (results_ts STOP conditions as flags)\n");
}

int main(int argc, char **argv) {
    printf(" --- Starting DECIPHeR rainfall-runoff modelling (B200 ensemble path) ---\n");
    std::string workdir = ".", auto_file;
    int ids[3] = {0, 0, 0};
    long members = -1;
    bool write_text = true, flow_bin = false;
    std::string member_list;
    int flags = 0, n_gpus = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-auto" && i + 1 < argc) auto_file = argv[++i];
        else if (a == "-ids" && i + 3 < argc) { ids[0] = atoi(argv[++i]); ids[1] = atoi(argv[++i]); ids[2] = atoi(argv[++i]); }
        else if (a == "-dir" && i + 1 < argc) workdir = argv[++i];
        else if (a == "-members" && i + 1 < argc) members = atol(argv[++i]);
        else if (a == "-no-flow") write_text = false;
        else if (a == "-write-members" && i + 1 < argc) member_list = argv[++i];
        else if (a == "-flow-bin") flow_bin = true;
        else if (a == "-gpus" && i + 1 < argc) n_gpus = atoi(argv[++i]);
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

    // members whose series are wanted on the host: all (reference behaviour), none (-no-flow) or the -write-members list
    std::vector<long> selected;                       // 0-based, ascending, unique
    if (!member_list.empty()) {
        std::vector<int64_t> tmp((size_t)numsim);
        const int64_t n_sel = dcph_parse_member_list(member_list.c_str(), numsim, tmp.data(), numsim);
        if (n_sel < 0) { fprintf(stderr, "%s\n", dcph_last_error()); return 2; }
        selected.assign(tmp.begin(), tmp.begin() + n_sel);
        write_text = true;
    } else if (write_text) {
        for (long m = 0; m < numsim; ++m) selected.push_back(m);
    }

    if (dcp_device_count() < 1) { fprintf(stderr, "no CUDA device: this build has no CPU fallback\n"); return 1; }
    if (n_gpus < 1 || n_gpus > DCP_MAX_RANKS) { fprintf(stderr, "-gpus must be in 1..%d\n", DCP_MAX_RANKS); return 2; }
    dcp_catchment *cat = nullptr;
    if (dcp_catchment_create(d, &cat) != DCP_OK) { fprintf(stderr, "dcp_catchment_create: %s\n", dcp_last_error()); return 1; }

    const size_t npar = (size_t)d->npt * 7;
    std::vector<double> mcpar(npar * numsim), perf((size_t)DCP_NMEASURE * d->nriv * numsim),
        totals(n_gpus > 1 ? 0 : (size_t)3 * d->nriv * numsim), diag((size_t)3 * numsim);
    std::vector<int32_t> status(numsim);
    dcph_genpar(M, numsim, mcpar.data());
    if (log) fprintf(log, " \n Looping through simulations...\n");

    dcp_run_opts opts{};
    opts.struct_size = sizeof(opts);
    opts.flags = flags;
    const auto t0 = std::chrono::steady_clock::now();
    // (1) every member: measures, reach totals, diagnostics, status -- nothing of length nstep leaves the device
    if (n_gpus > 1) {
        dcp_comm *comm = nullptr;
        if (dcp_comm_init(n_gpus, nullptr, &comm) != DCP_OK) { fprintf(stderr, "dcp_comm_init: %s\n", dcp_last_error()); return 1; }
        dcp_multi_stats ms{};
        // reach totals are only printed for the selected members (re-run below), so the multi-device pass gathers measures only
        if (dcp_run_ensemble_multi(comm, d, numsim, mcpar.data(), &opts, nullptr, perf.data(), nullptr, diag.data(), status.data(), &ms) != DCP_OK) {
            fprintf(stderr, "dcp_run_ensemble_multi: %s\n", dcp_last_error());
            return 1;
        }
        printf(" %d devices%s: slowest rank %.1f ms busy, mean %.1f ms, gather %.2f ms\n", ms.n_ranks, ms.used_nccl ? " (NCCL gather)" : "",
               ms.ms_busy_max, ms.ms_busy_sum / ms.n_ranks, ms.ms_gather);
        dcp_comm_destroy(comm);
    } else {
        const int rc = dcp_run_ensemble(cat, 1, numsim, mcpar.data(), &opts, nullptr, perf.data(), totals.data(), diag.data(), status.data());
        if (rc != DCP_OK) { fprintf(stderr, "dcp_run_ensemble: %s\n", dcp_last_error()); return 1; }
    }
    // (2) the selected members again, in chunks that fit in host memory, with their flow series: text files in the
    //     reference's formats and / or the binary flow table.  Members are independent, so re-running a subset gives
    //     the same bits as the ensemble pass.
    FILE *fbin = nullptr;
    if (flow_bin && !selected.empty()) {
        const std::string fn = std::string(dcph_out_prefix(M)) + "ensemble_flow.bin";
        fbin = fopen(fn.c_str(), "wb");
        if (!fbin) { fprintf(stderr, "cannot open %s\n", fn.c_str()); return 1; }
        const char magic[8] = {'D', 'C', 'P', 'F', 'L', 'O', 'W', '1'};
        const int64_t hdr[3] = {(int64_t)selected.size(), d->nriv, d->nstep};
        fwrite(magic, 1, 8, fbin); fwrite(hdr, 8, 3, fbin);
        for (long m : selected) { const int64_t id = m + 1; fwrite(&id, 8, 1, fbin); }       // member ids, then q[nriv x nstep] each
    }
    const long chunk = std::max<long>(1, (long)(2e9 / (8.0 * d->nriv * d->nstep)));
    std::vector<double> q, mc_sel, perf_sel, tot_sel, diag_sel;
    std::vector<int32_t> st_sel;
    for (size_t s0 = 0; s0 < selected.size(); s0 += chunk) {
        const long n = (long)std::min<size_t>(chunk, selected.size() - s0);
        q.resize((size_t)d->nriv * d->nstep * n); mc_sel.resize(npar * n); perf_sel.resize((size_t)DCP_NMEASURE * d->nriv * n);
        tot_sel.resize((size_t)3 * d->nriv * n); diag_sel.resize((size_t)3 * n); st_sel.resize(n);
        // parameter sets BEFORE the clip are needed as input: regenerate the table once (cheap, bit-exact RANMAR)
        static std::vector<double> mc_in;
        if (mc_in.empty()) { mc_in.resize(npar * numsim); dcph_genpar(M, numsim, mc_in.data()); }
        for (long k = 0; k < n; ++k) std::memcpy(&mc_sel[npar * k], &mc_in[npar * selected[s0 + k]], npar * 8);
        const int rc = dcp_run_ensemble(cat, selected[s0] + 1, n, mc_sel.data(), &opts, q.data(), perf_sel.data(), tot_sel.data(),
                                        diag_sel.data(), st_sel.data());
        if (rc != DCP_OK) { fprintf(stderr, "dcp_run_ensemble (selected members): %s\n", dcp_last_error()); return 1; }
        for (long k = 0; k < n; ++k) {
            const long m = selected[s0 + k];
            printf(" Running sim %12ld\n", m + 1);
            if (dcph_write_member(M, m + 1, &mc_sel[npar * k], &q[(size_t)d->nriv * d->nstep * k], &tot_sel[(size_t)3 * d->nriv * k],
                                  &diag_sel[(size_t)3 * k], st_sel[k], 0) != DCP_OK) {
                fprintf(stderr, "write_output: %s\n", dcph_last_error());
                return 1;
            }
            if (fbin) fwrite(&q[(size_t)d->nriv * d->nstep * k], 8, (size_t)d->nriv * d->nstep, fbin);
        }
    }
    if (fbin) fclose(fbin);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // binary table: the sampled parameter sets (after the srinit clip) next to their measures
    const std::string table = std::string(dcph_out_prefix(M)) + "ensemble_perf.bin";
    if (FILE *f = fopen(table.c_str(), "wb")) {
        const char magic[8] = {'D', 'C', 'P', 'P', 'E', 'R', 'F', '2'};
        const int64_t hdr[4] = {numsim, d->nriv, DCP_NMEASURE, d->npt};
        fwrite(magic, 1, 8, f);
        fwrite(hdr, 8, 4, f);
        fwrite(mcpar.data(), 8, mcpar.size(), f);
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
