// decipher_route_main.cpp -- the standalone river router with the reference executable's interface
// (DTA/route_run_standalone.f90:72-189):
//
//     decipher_route_run -river_data <xxx_river_data.txt> -tree <xxx_flow_conn.txt> -timeseries_in <grid.txt> [-velocity v] [-velocities v1,v2,...]
//
// reads the routing tree (xxx_flow_point.txt is found beside the tree file), the river cells and the flow series (ASCII grid: one
// row per timestep, one column per node, or one column less when the first node is an ungauged sea outlet), builds the
// time-delay histograms for the constant velocity (reference default 0.133 metres per timestep) and writes
// <grid minus extension>_routed.txt like write_numeric_list (6 decimals).  -velocities routes the SAME series once per velocity
// in one dcp_route_timeseries call (what the GPU is for) and writes <...>_routed_<k>.txt per velocity.
#include "dcp_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char **argv) {
    printf(" --- route_run_standalone (B200) ---\n");
    std::string river_data, tree, series;
    std::vector<double> vel;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-river_data" && i + 1 < argc) river_data = argv[++i];
        else if (a == "-tree" && i + 1 < argc) tree = argv[++i];
        else if (a == "-timeseries_in" && i + 1 < argc) series = argv[++i];
        else if (a == "-velocity" && i + 1 < argc) vel.assign(1, atof(argv[++i]));
        else if (a == "-velocities" && i + 1 < argc) {
            vel.clear();
            for (char *tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ",")) vel.push_back(atof(tok));
        }
    }
    if (river_data.empty() || tree.empty() || series.empty()) {
        printf("  command options \n -river_data <xxx_river_data.txt> river data (from route_river_file.e)\n"
               " -tree <xxx_flow_conn.txt> the routing tree file\n         note: xxx_flow_point.txt will also be read \n -timeseries_in\n"
               " -velocity <metres per timestep> (default 0.133)   -velocities v1,v2,... (one routed file per velocity)\n");
        return 2;
    }
    if (vel.empty()) vel.assign(1, 0.133);                 // tdh%v_param(1) = 0.133d0 (route_run_standalone.f90:61)
    dcph_model *M = nullptr;
    if (dcph_load_router(river_data.c_str(), tree.c_str(), series.c_str(), &M) != DCP_OK) {
        fprintf(stderr, "%s\n", dcph_last_error());
        return 1;
    }
    const dcp_catchment_desc *d = dcph_desc(M);
    printf(" %d nodes, %d flow columns, %d timesteps, %zu velocit%s\n", d->nnode, d->nriv, d->nstep, vel.size(), vel.size() == 1 ? "y" : "ies");
    if (dcp_device_count() < 1) { fprintf(stderr, "no CUDA device: this build has no CPU fallback\n"); return 1; }
    dcp_catchment *cat = nullptr;
    if (dcp_catchment_create(d, &cat) != DCP_OK) { fprintf(stderr, "dcp_catchment_create: %s\n", dcp_last_error()); return 1; }
    const size_t per = (size_t)d->nriv * d->nstep, n = vel.size();
    std::vector<double> inflow(per * n), routed(per * n);
    for (size_t s = 0; s < n; ++s) std::memcpy(&inflow[per * s], dcph_router_inflow(M), per * 8);
    std::vector<int32_t> status(n);
    if (dcp_route_timeseries(cat, (int64_t)n, vel.data(), inflow.data(), routed.data(), status.data(), 0) != DCP_OK) {
        fprintf(stderr, "dcp_route_timeseries: %s\n", dcp_last_error());
        return 1;
    }
    const std::string stem = series.size() > 4 ? series.substr(0, series.size() - 4) : series;      // route_run_standalone.f90:177
    for (size_t s = 0; s < n; ++s) {
        const std::string out = stem + (n == 1 ? std::string("_routed.txt") : "_routed_" + std::to_string(s + 1) + ".txt");
        printf(" write output: %s\n Velocity %g%s\n", out.c_str(), vel[s], (status[s] & DCP_ST_HIST_DROPPED) ? "  (histogram cells dropped)" : "");
        if (dcph_write_routed(M, &routed[per * s], out.c_str(), 6, 0) != DCP_OK) { fprintf(stderr, "%s\n", dcph_last_error()); return 1; }
    }
    dcp_catchment_destroy(cat);
    dcph_free(M);
    return 0;
}
