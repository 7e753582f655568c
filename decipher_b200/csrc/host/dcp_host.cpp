// dcp_host.cpp — host side of the drop-in (see dcp_host.h).  Reads the reference's text formats with
// Fortran list-directed semantics, samples parameters with RANMAR, writes .res / .flow files.
#include "dcp_host.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct HostError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] void die(const char *fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw HostError(buf);
}

// ------------------------------------------------------------------------------------------------
// A Fortran sequential formatted unit: records = lines.  Each READ starts at a fresh record; a
// list-directed READ pulls blank/comma/tab separated tokens across as many records as it needs and
// discards the rest of the last record it touched.
// ------------------------------------------------------------------------------------------------
class Unit {
public:
    explicit Unit(const std::string &path) : path_(path) {
        std::ifstream f(path);
        if (!f) die("cannot open '%s'", path.c_str());
        std::string line;
        while (std::getline(f, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            lines_.push_back(line);
        }
    }
    void skip_record() { next_record("skip"); }                 // read(u,*) with no list / read(u,'(A)')
    std::string read_record() { return next_record("record"); }
    void begin_list() { col_ = std::string::npos; }
    // next list-directed token of the current READ
    std::string token() {
        for (;;) {
            if (col_ == std::string::npos) {
                if (rec_ >= lines_.size()) die("%s: end of file in list-directed read", path_.c_str());
                cur_ = lines_[rec_++];
                col_ = 0;
            }
            while (col_ < cur_.size() && is_sep(cur_[col_])) ++col_;
            if (col_ >= cur_.size()) { col_ = std::string::npos; continue; }
            size_t b = col_;
            while (col_ < cur_.size() && !is_sep(cur_[col_])) ++col_;
            return cur_.substr(b, col_ - b);
        }
    }
    void end_list() { col_ = std::string::npos; }
    double real() {
        std::string t = token();
        for (auto &ch : t) if (ch == 'd' || ch == 'D') ch = 'e';
        char *e = nullptr;
        const double v = std::strtod(t.c_str(), &e);
        if (e == t.c_str()) die("%s: expected a number, found '%s' (record %zu)", path_.c_str(), t.c_str(), rec_);
        return v;
    }
    int integer() {
        const std::string t = token();
        char *e = nullptr;
        const long v = std::strtol(t.c_str(), &e, 10);
        if (e == t.c_str()) die("%s: expected an integer, found '%s' (record %zu)", path_.c_str(), t.c_str(), rec_);
        return (int)v;
    }
    bool at_end() const { return rec_ >= lines_.size(); }
    std::string peek_record() const { return rec_ < lines_.size() ? lines_[rec_] : std::string(); }   // next record, not consumed
    size_t record_index() const { return rec_; }
    void rewind() { rec_ = 0; col_ = std::string::npos; }
    // counts the records from the cursor that hold at least one token (read(fid,*,iostat=io) tmp loop)
    size_t count_token_records() {
        size_t n = 0;
        for (size_t r = rec_; r < lines_.size(); ++r)
            if (lines_[r].find_first_not_of(" \t,") != std::string::npos) ++n;
        return n;
    }
    const std::string &path() const { return path_; }

private:
    static bool is_sep(char c) { return c == ' ' || c == '\t' || c == ','; }
    std::string next_record(const char *what) {
        col_ = std::string::npos;
        if (rec_ >= lines_.size()) die("%s: end of file (%s)", path_.c_str(), what);
        return lines_[rec_++];
    }
    std::string path_;
    std::vector<std::string> lines_;
    size_t rec_ = 0, col_ = std::string::npos;
    std::string cur_;
};

std::string rtrim(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t')) s.pop_back();
    return s;
}

// format(a13,1X,a700): characters 15.. of the record (dyna_project.f90:330)
std::string field_a700(const std::string &rec) { return rtrim(rec.size() > 14 ? rec.substr(14, 700) : std::string()); }
// format(a13,1X,i4)
int field_i4(const std::string &rec, const std::string &path) {
    const std::string f = rec.size() > 14 ? rec.substr(14, 4) : std::string();
    char *e = nullptr;
    const long v = std::strtol(f.c_str(), &e, 10);
    if (e == f.c_str()) die("%s: expected an integer in columns 15-18 of '%s'", path.c_str(), rec.c_str());
    return (int)v;
}

double seqsum(const std::vector<double> &v) {
    double s = 0;
    for (double x : v) s += x;
    return s;
}

// ------------------------------------------------------------------------------------------------
// RANMAR (RRMODEL/dyna_random.f90)
// ------------------------------------------------------------------------------------------------
struct Ranmar {
    double cu[98], cc, cd, cm;
    int i97, j97;
    void ranin(int ij, int kl) {                                  // :111-133
        int i = (ij / 177) % 177 + 2, j = ij % 177 + 2, k = (kl / 169) % 178 + 1, l = kl % 169;
        for (int ii = 1; ii <= 97; ++ii) {
            double s = 0.0, t = 0.5;
            for (int jj = 1; jj <= 24; ++jj) {
                const int m = (((i * j) % 179) * k) % 179;
                i = j; j = k; k = m;
                l = (53 * l + 1) % 169;
                if ((l * m) % 64 >= 32) s += t;
                t *= 0.5;
            }
            cu[ii] = s;
        }
        cc = 362436.0 / 16777216.0; cd = 7654321.0 / 16777216.0; cm = 16777213.0 / 16777216.0;
        i97 = 97; j97 = 33;
    }
    double next() {                                               // :51-64
        double uni = cu[i97] - cu[j97];
        if (uni < 0.0) uni += 1.0;
        cu[i97] = uni;
        if (--i97 == 0) i97 = 97;
        if (--j97 == 0) j97 = 97;
        cc -= cd;
        if (cc < 0.0) cc += cm;
        uni -= cc;
        if (uni < 0.0) uni += 1.0;
        return uni;
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
struct dcph_model {
    dcp_catchment_desc d{};
    // storage behind the descriptor
    std::vector<double> ac, st, mslp, wtrans, rivers, sum_ac_riv, qobs_start, frac_sub_area, river_data, rain, pet, qobs;
    std::vector<int32_t> ippt, ipet, ipar, ipriv, ims, ims_rz, ims_uz, ntrans, itrans, node_gauge_id, node_type,
        uptree_count, uptree_index, node_to_flow_mapping;
    // Monte-Carlo setup
    int numsim = 0, seed_1 = 0, seed_2 = 0, ntt = 1;
    double wt = 0, acc = 0;
    std::vector<double> ll, ul;          // [npt x 7]
    std::string out_prefix, log_path, proj_dir;
    std::vector<std::string> log;
    // standalone router (dcph_load_router): the input series in the ABI layout [nriv x nstep] and the output column headers
    std::vector<double> router_inflow;
    std::vector<std::string> router_headers;
};

namespace {

const char *PM[7] = {"szm", "lnto", "srmax", "srinit", "chv", "td", "smax"};

void read_tread_dyna(dcph_model &M, const std::string &meta_file, const std::string &flux_file) {
    // ---- HRU meta file (dyna_tread_dyna.f90:46-100) ----
    Unit u17(meta_file);
    u17.begin_list(); u17.token(); u17.end_list();                       // read(17,*) tmp
    u17.begin_list(); u17.token(); const int nac = u17.integer(); u17.end_list();
    u17.begin_list(); u17.token(); const int num_cols = u17.integer(); u17.end_list();
    u17.begin_list(); u17.token(); u17.end_list();
    if (nac < 1 || num_cols < 5) die("%s: HRUs=%d NUM_COLS=%d", meta_file.c_str(), nac, num_cols);
    std::vector<std::string> cols(num_cols);
    u17.begin_list();
    for (auto &c : cols) c = u17.token();
    u17.end_list();
    std::vector<double> meta((size_t)nac * num_cols);
    for (int i = 0; i < nac; ++i) {
        u17.begin_list();
        for (int j = 0; j < num_cols; ++j) meta[(size_t)j * nac + i] = u17.real();
        u17.end_list();
    }
    M.ac.resize(nac); M.st.resize(nac); M.mslp.resize(nac);
    M.ipriv.resize(nac); M.ippt.assign(nac, 1); M.ipet.assign(nac, 1); M.ipar.assign(nac, 1); M.ims.assign(nac, 1);
    double cells = 0;
    for (int i = 0; i < nac; ++i) cells += meta[(size_t)1 * nac + i];     // sum(hru_meta(:,2))
    for (int i = 0; i < nac; ++i) {                                       // :67-72
        M.ac[i] = meta[(size_t)1 * nac + i] / cells;
        M.st[i] = meta[(size_t)2 * nac + i];
        M.mslp[i] = meta[(size_t)3 * nac + i];
        M.ipriv[i] = (int)meta[(size_t)4 * nac + i];
    }
    for (int c = 0; c < num_cols; ++c) {                                  // :80-100
        std::vector<int32_t> *dst = nullptr;
        if (cols[c] == "RAIN_ID") dst = &M.ippt;
        else if (cols[c] == "PET_ID") dst = &M.ipet;
        else if (cols[c] == "PARAM_ID") dst = &M.ipar;
        else if (cols[c] == "MODELSTRUCT_ID") dst = &M.ims;
        if (dst) for (int i = 0; i < nac; ++i) (*dst)[i] = (int)meta[(size_t)c * nac + i];
    }
    // ---- HRU flux file (:109-170) ----
    Unit u13(flux_file);
    u13.begin_list(); u13.token(); u13.end_list();
    u13.begin_list(); u13.token(); const int nac2 = u13.integer(); u13.end_list();
    u13.begin_list(); u13.token(); u13.real(); u13.end_list();            // catchment area (unused downstream)
    u13.begin_list(); u13.token(); const int num_class = u13.integer(); u13.end_list();
    u13.begin_list(); u13.token(); const int nriv = u13.integer(); u13.end_list();
    if (nac2 != nac) die("%s: %d HRUs but the meta file has %d", flux_file.c_str(), nac2, nac);
    if (nriv < 1) die("%s: Gauges = %d", flux_file.c_str(), nriv);
    for (int i = 0; i < num_class; ++i) { u13.begin_list(); u13.token(); u13.end_list(); }   // :122-124
    M.ntrans.resize(nac);
    M.itrans.clear(); M.wtrans.clear();
    for (int ia = 0; ia < nac; ++ia) {
        u13.begin_list();                                                 // 'HRU' i 'Sharing' 'to' n 'HRUs'
        u13.token(); u13.integer(); u13.token(); u13.token();
        const int nt = u13.integer();
        u13.token();
        u13.end_list();
        if (nt < 0) die("%s: HRU %d shares to %d HRUs", flux_file.c_str(), ia + 1, nt);
        M.ntrans[ia] = nt;
        for (int iw = 0; iw < nt; ++iw) {
            u13.begin_list();
            M.itrans.push_back(u13.integer());
            M.wtrans.push_back(u13.real());
            u13.end_list();
        }
        u13.begin_list(); u13.token(); M.wtrans.push_back(u13.real()); u13.end_list();       // RIVS w
    }
    M.rivers.assign((size_t)nac * nriv, 0.0);
    M.sum_ac_riv.assign(nriv, 0.0);
    u13.begin_list(); u13.token(); u13.end_list();                        // header line
    for (int ia = 0; ia < nac; ++ia) {
        u13.begin_list();
        u13.token(); u13.integer(); u13.token(); u13.token();
        const int nr = u13.integer();
        u13.token();
        u13.end_list();
        for (int iw = 0; iw < nr; ++iw) {
            u13.begin_list();
            const int rid = u13.integer();
            const double share = u13.real();
            u13.end_list();
            if (rid < 1 || rid > nriv) die("%s: HRU %d river id %d out of range", flux_file.c_str(), ia + 1, rid);
            M.rivers[(size_t)(rid - 1) * nac + ia] = share;
        }
        if (M.ipriv[ia] < 1 || M.ipriv[ia] > nriv) die("%s: HRU %d GAUGE_ID %d out of range", meta_file.c_str(), ia + 1, M.ipriv[ia]);
        M.sum_ac_riv[M.ipriv[ia] - 1] += M.ac[ia];                        // :166-167 (before the ac renormalisation)
    }
    // weights sum to one (:173-184)
    size_t off = 0;
    for (int ia = 0; ia < nac; ++ia) {
        const int n = M.ntrans[ia] + 1;
        double s = 0;
        for (int iw = 0; iw < n; ++iw) s += M.wtrans[off + iw];
        if (s != 1.0) for (int iw = 0; iw < n; ++iw) M.wtrans[off + iw] = M.wtrans[off + iw] / s;
        off += n;
    }
    // ac renormalised inside a loop that recomputes sum(ac) every iteration (:186-190)
    if (seqsum(M.ac) != 1.0)
        for (int ia = 0; ia < nac; ++ia) M.ac[ia] = M.ac[ia] / seqsum(M.ac);
    const double s_ac = seqsum(M.ac);
    for (int i = 0; i < nriv; ++i) M.sum_ac_riv[i] = M.sum_ac_riv[i] / s_ac;                // :192-194
    M.d.nac = nac; M.d.nriv = nriv;
}

// reads a forcing-type file: 3 skipped records, header "dt nstep n [flow_err]", 2 skipped, then rows (dyna_inputs.f90)
void read_series(const std::string &file, bool has_err, double &dt, int &nstep, int &ncol, double &flow_err,
                 std::vector<double> &data /*[ncol x nstep]*/) {
    Unit u(file);
    u.skip_record(); u.skip_record(); u.skip_record();
    u.begin_list();
    dt = u.real(); nstep = u.integer(); ncol = u.integer();
    if (has_err) flow_err = u.real();
    u.end_list();
    u.skip_record(); u.skip_record();
    if (nstep < 1 || ncol < 1) die("%s: nstep=%d columns=%d", file.c_str(), nstep, ncol);
    data.assign((size_t)ncol * nstep, 0.0);
    for (int it = 0; it < nstep; ++it) {
        u.begin_list();
        for (int k = 0; k < 4; ++k) u.integer();                          // Y M D H
        for (int c = 0; c < ncol; ++c) data[(size_t)it * ncol + c] = u.real();
        u.end_list();
    }
}

void read_inputs(dcph_model &M, const std::string &flow_file, const std::string &rain_file, const std::string &pet_file) {
    double dt = 0, ferr = 0;
    int nstep_flow = 0, nriv2 = 0;
    read_series(flow_file, true, dt, nstep_flow, nriv2, ferr, M.qobs);
    if (nriv2 != M.d.nriv)                                                // dyna_inputs.f90:65-70
        die("The number of gauges in your input (%d) and HRU files (%d) do not match", nriv2, M.d.nriv);
    const int nriv = M.d.nriv;
    M.qobs_start.assign(nriv, 0.0);
    for (int r = 0; r < nriv; ++r) M.qobs_start[r] = M.qobs[r];           // first row (:91-94)
    for (int r = 0; r < nriv; ++r) {                                      // :101-113
        if (M.qobs_start[r] == ferr) {
            int n_err = 0;
            for (int t = 0; t < nstep_flow; ++t) if (M.qobs[(size_t)t * nriv + r] == ferr) ++n_err;
            if (n_err == nstep_flow) {
                M.qobs_start[r] = (double)(0.001f / 24.0f) * dt;          // default-real constant expression
            } else {
                double s = 0;
                for (int t = 0; t < nstep_flow; ++t) { const double v = M.qobs[(size_t)t * nriv + r]; if (v != ferr) s += v; }
                M.qobs_start[r] = s / (nstep_flow - n_err);
            }
        }
    }
    double ferr_unused = 0;
    int nstep_ppt = 0, nstep_pet = 0, ngp = 0, nge = 0;
    read_series(rain_file, false, dt, nstep_ppt, ngp, ferr_unused, M.rain);
    read_series(pet_file, false, dt, nstep_pet, nge, ferr_unused, M.pet);
    if (nstep_ppt != nstep_pet)                                           // :166-174
        die("The number of timesteps in your precipitation (%d) and PET (%d) time series MUST match", nstep_ppt, nstep_pet);
    M.d.nstep = nstep_ppt; M.d.ngp = ngp; M.d.nge = nge; M.d.dt = dt;     // dt of the last file read (PET)
    M.d.nstep_obs = nstep_flow; M.d.flow_err = ferr;
}

void read_mc_setup(dcph_model &M, const std::string &file) {
    Unit u(file);                                                          // dyna_mc_setup.f90:81-110
    u.begin_list(); M.numsim = u.integer(); u.end_list();
    u.begin_list(); M.seed_1 = u.integer(); M.seed_2 = u.integer(); u.end_list();
    u.begin_list(); M.ntt = u.integer(); M.wt = u.real(); M.acc = u.real(); u.end_list();
    u.begin_list(); const int npt = u.integer(); const int nnames = u.integer(); u.end_list();
    u.skip_record();
    int max_ipar = 0;
    for (int v : M.ipar) max_ipar = std::max(max_ipar, v);
    if (max_ipar != npt)                                                  // :89-94
        die("Number of parameter types must match the number given from the HRU analysis (file %d, HRU meta %d)", npt, max_ipar);
    std::vector<std::string> names(nnames + 1);
    u.begin_list();
    for (auto &n : names) n = u.token();
    u.end_list();
    std::vector<double> values((size_t)npt * nnames);
    for (int i = 0; i < npt; ++i) {
        u.begin_list();
        u.integer();
        for (int j = 0; j < nnames; ++j) values[(size_t)j * npt + i] = u.real();
        u.end_list();
    }
    M.ll.assign((size_t)npt * 7, 0.0); M.ul.assign((size_t)npt * 7, 0.0);
    double declared[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 1; i <= nnames; ++i) {                                   // :116-155
        bool found = false;
        for (int j = 0; j < 7; ++j) {
            const std::string base = PM[j];
            auto col = [&](std::vector<double> &dst) { for (int l = 0; l < npt; ++l) dst[(size_t)j * npt + l] = values[(size_t)(i - 1) * npt + l]; };
            if (names[i] == base + "_def") { col(M.ll); col(M.ul); declared[j] = 1; found = true; }
            if (names[i] == base + "_ll") { col(M.ll); declared[j] += 0.5; found = true; }
            if (names[i] == base + "_ul") { col(M.ul); declared[j] += 0.5; found = true; }
        }
        if (!found) die("ERROR: %s has been specified in param file but does not exist", names[i].c_str());
    }
    double tot = 0;
    for (double x : declared) tot += x;
    if (tot < 7) {                                                        // :159-169
        std::string miss;
        for (int j = 0; j < 7; ++j) if (declared[j] < 1) miss += std::string(" ") + PM[j];
        die("ERROR: Parameters missing from the parameter file:%s", miss.c_str());
    }
    if (M.ntt < 1) die("NTT must be >= 1");
    M.d.npt = npt; M.d.ntt = M.ntt;
}

void read_modelstruct(dcph_model &M, const std::string &file) {
    Unit u(file);                                                          // dyna_modelstruct_setup.f90:30-111
    u.skip_record(); u.skip_record();
    u.begin_list(); const int ntypes = u.integer(); const int nnames = u.integer(); u.end_list();
    int max_ims = 0;
    for (int v : M.ims) max_ims = std::max(max_ims, v);
    if (max_ims != ntypes) die("Number of model structure types must match the number given from the HRU analysis (file %d, HRU meta %d)", ntypes, max_ims);
    std::vector<std::string> names(nnames + 1);
    u.begin_list();
    for (auto &n : names) n = u.token();
    u.end_list();
    std::vector<int> values((size_t)ntypes * nnames);
    for (int i = 0; i < ntypes; ++i) {
        u.begin_list();
        u.integer();
        for (int j = 0; j < nnames; ++j) values[(size_t)j * ntypes + i] = u.integer();
        u.end_list();
    }
    const int nac = M.d.nac;
    M.ims_rz.assign(nac, 1); M.ims_uz.assign(nac, 1);
    for (int i = 1; i <= nnames; ++i) {
        std::vector<int32_t> *dst = names[i] == "rootzone" ? &M.ims_rz : (names[i] == "unsatzone" ? &M.ims_uz : nullptr);
        if (!dst) continue;                                               // unknown names only print a message upstream
        for (int j = 0; j < nac; ++j) {
            if (M.ims[j] < 1 || M.ims[j] > ntypes) die("MODELSTRUCT_ID %d of HRU %d out of range", M.ims[j], j + 1);
            (*dst)[j] = values[(size_t)(i - 1) * ntypes + (M.ims[j] - 1)];
        }
    }
}

// read_numeric_list_fid (DTA/dta_utility.f90:352-399): header rows skipped, then every record with a token is a row
void read_numeric_list(const std::string &file, int columns, int header_rows, std::vector<double> &data, int &nrow) {
    Unit u(file);
    for (int i = 0; i < header_rows; ++i) u.skip_record();
    nrow = (int)u.count_token_records();
    data.assign((size_t)nrow * columns, 0.0);                             // column-major [nrow x columns]
    for (int i = 0; i < nrow; ++i) {
        u.begin_list();
        for (int c = 0; c < columns; ++c) data[(size_t)c * nrow + i] = u.real();
        u.end_list();
    }
}

void riv_full_upstream(int node, const std::vector<std::vector<int>> &ups, std::vector<int32_t> &acc) {
    for (int u : ups[node]) acc.push_back(u + 1);                         // direct links first (riv_tree_node.f90:97-103)
    for (int u : ups[node]) riv_full_upstream(u, ups, acc);               // then recurse (:105-108)
}

void read_routingdata(dcph_model &M, const std::string &flow_conn, const std::string &riv_data, const std::string &flow_point) {
    std::vector<double> fp, fc;
    int n_fp = 0, n_fc = 0, ncell = 0;
    read_numeric_list(flow_point, 6, 1, fp, n_fp);                        // riv_tree_read_dyna (:605-606)
    read_numeric_list(flow_conn, 10, 1, fc, n_fc);
    read_numeric_list(riv_data, 8, 1, M.river_data, ncell);               // dta_route_processing.f90:110
    if (n_fp < 1 || ncell < 1) die("routing files are empty");
    const int nnode = n_fp;
    M.node_gauge_id.resize(nnode); M.node_type.resize(nnode);
    for (int i = 0; i < nnode; ++i) {                                     // riv_tree_read_impl (:492-512)
        M.node_gauge_id[i] = (int)std::lround(fp[(size_t)0 * nnode + i]);
        M.node_type[i] = (int)std::lround(fp[(size_t)3 * nnode + i]);
    }
    std::vector<int> down(nnode, 0);
    for (int i = 0; i < n_fc; ++i) {                                      // :514-547
        const int node_id = (int)std::lround(fc[(size_t)0 * n_fc + i]), down_id = (int)std::lround(fc[(size_t)4 * n_fc + i]);
        int idx = -1;
        for (int j = 0; j < nnode; ++j) if (M.node_gauge_id[j] == node_id) { idx = j; break; }
        if (idx < 0) die("Point file and flow file do not match (node %d)", node_id);
        for (int j = 0; j < nnode; ++j) if (M.node_gauge_id[j] == down_id) { down[idx] = j + 1; break; }
    }
    std::vector<std::vector<int>> ups(nnode);                             // build_upstream_from_downstream (:563-590)
    for (int i = 0; i < nnode; ++i) if (down[i] > 0) ups[down[i] - 1].push_back(i);
    M.uptree_count.resize(nnode); M.uptree_index.clear();
    std::vector<std::vector<int32_t>> tree(nnode);
    for (int i = 0; i < nnode; ++i) {
        riv_full_upstream(i, ups, tree[i]);
        M.uptree_count[i] = (int)tree[i].size();
        M.uptree_index.insert(M.uptree_index.end(), tree[i].begin(), tree[i].end());
    }
    // frac_sub_area (dyna_read_routingdata.f90:52-89): own HRUs ascending, then each upstream node's
    const int nac = M.d.nac;
    M.frac_sub_area.assign(nnode, 0.0);
    for (int i = 0; i < nnode; ++i) {
        double f = 0;
        for (int z = 0; z < nac; ++z) if (M.ipriv[z] == i + 1) f = M.ac[z] + f;
        for (int k : tree[i])
            for (int z = 0; z < nac; ++z) if (k == M.ipriv[z]) f = M.ac[z] + f;
        M.frac_sub_area[i] = f;
    }
    M.node_to_flow_mapping.resize(M.d.nriv);
    for (int i = 0; i < M.d.nriv; ++i) M.node_to_flow_mapping[i] = i + 1;  // :100-102
    M.d.nnode = nnode; M.d.ncell = ncell;
}

void finish_desc(dcph_model &M) {
    dcp_catchment_desc &d = M.d;
    d.struct_size = sizeof(dcp_catchment_desc);
    d.ac = M.ac.data(); d.st = M.st.data(); d.mslp = M.mslp.data();
    d.ippt = M.ippt.data(); d.ipet = M.ipet.data(); d.ipar = M.ipar.data(); d.ipriv = M.ipriv.data();
    d.ims_rz = M.ims_rz.data(); d.ims_uz = M.ims_uz.data();
    d.ntrans = M.ntrans.data(); d.itrans = M.itrans.data(); d.wtrans = M.wtrans.data();
    d.rivers = M.rivers.data(); d.sum_ac_riv = M.sum_ac_riv.data(); d.qobs_start = M.qobs_start.data();
    d.node_gauge_id = M.node_gauge_id.data(); d.node_type = M.node_type.data();
    d.uptree_count = M.uptree_count.data(); d.uptree_index = M.uptree_index.data();
    d.frac_sub_area = M.frac_sub_area.data(); d.river_data = M.river_data.data();
    d.node_to_flow_mapping = M.node_to_flow_mapping.data();
    d.rain = M.rain.data(); d.pet = M.pet.data(); d.qobs = M.qobs.data();
    d.t_warm = 0;
    for (int i = 0; i < d.nac; ++i) {
        if (M.ippt[i] < 1 || M.ippt[i] > d.ngp) die("HRU %d: RAIN_ID %d but the rain file has %d columns", i + 1, M.ippt[i], d.ngp);
        if (M.ipet[i] < 1 || M.ipet[i] > d.nge) die("HRU %d: PET_ID %d but the PET file has %d columns", i + 1, M.ipet[i], d.nge);
    }
}

// ------------------------------------------------------------------------------------------------
// Fortran edit descriptors used by write_output
// ------------------------------------------------------------------------------------------------
std::string fmt_F(double x, int w, int dgt) {            // Fw.d
    char b[512];
    snprintf(b, sizeof(b), "%.*f", dgt, x);
    std::string s = b;
    if ((int)s.size() > w && s.rfind("0.", 0) == 0) s = s.substr(1);          // optional leading zero dropped to fit
    if ((int)s.size() > w && s.rfind("-0.", 0) == 0) s = "-" + s.substr(2);
    if ((int)s.size() > w) return std::string(w, '*');
    return std::string(w - s.size(), ' ') + s;
}
std::string fmt_F0(double x, int dgt, bool leading_zero) {   // F0.d (minimal width)
    char b[512];
    snprintf(b, sizeof(b), "%.*f", dgt, x);
    std::string s = b;
    if (!leading_zero) {
        if (s.rfind("0.", 0) == 0) s = s.substr(1);
        else if (s.rfind("-0.", 0) == 0) s = "-" + s.substr(2);
    }
    return s;
}
std::string fmt_A(const std::string &s20, int w) {          // Aw of a character(len=20) item
    std::string s = s20;
    s.resize(20, ' ');
    if (w <= 20) return s.substr(0, w);
    return std::string(w - 20, ' ') + s;
}
// gfortran list-directed real(8) (G25.17E3 rule): 25 columns.  0.1 <= |x| < 1e17 -> F editing with 17
// significant digits right-justified in 20 columns followed by 5 blanks, otherwise d.dddddddddddddddd E+ddd.
// The spacing of list-directed output is compiler specific; only the numbers are meaningful.
std::string fmt_list_real(double x) {
    char b[80];
    std::string s;
    if (!std::isfinite(x)) {
        s = std::isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity");
        return std::string(25 - s.size(), ' ') + s;
    }
    const double a = std::fabs(x);
    if (a == 0.0 || (a >= 0.1 && a < 1e17)) {
        int n = 0;                                            // digits before the decimal point
        if (a >= 1.0) n = (int)std::floor(std::log10(a)) + 1;
        snprintf(b, sizeof(b), "%.*f", 17 - std::max(n, 1) + (a < 1.0 && a != 0.0 ? 1 : 0), x);
        s = b;
        if (s.size() < 20) s = std::string(20 - s.size(), ' ') + s;
        return s + "     ";
    }
    snprintf(b, sizeof(b), "%.16E", x);
    s = b;
    const size_t e = s.find('E');
    std::string mant = s.substr(0, e), digits = s.substr(e + 2);
    while (digits.size() < 3) digits = "0" + digits;
    s = mant + "E" + s[e + 1] + digits;
    if (s.size() < 25) s = std::string(25 - s.size(), ' ') + s;
    return s;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char *dcph_last_error(void) { return g_err.c_str(); }

void dcph_ranmar(int ij, int kl, int64_t n_skip, int64_t n, double *out) {
    Ranmar r;
    r.ranin(ij, kl);
    for (int64_t i = 0; i < n_skip; ++i) r.next();
    for (int64_t i = 0; i < n; ++i) out[i] = r.next();
}

int dcph_read_auto_file(const char *path, int *id_project, int *id_cat, int *id_flow) {
    try {
        Unit u(path);
        u.begin_list(); *id_project = u.integer(); u.end_list();
        u.begin_list(); *id_cat = u.integer(); u.end_list();
        u.begin_list(); *id_flow = u.integer(); u.end_list();
        return DCP_OK;
    } catch (const std::exception &e) { g_err = e.what(); return DCP_ERR_INVALID; }
}

int dcph_load_project(const char *workdir, int id_project, int id_cat, int id_flow, dcph_model **out) {
    if (!out) return DCP_ERR_INVALID;
    *out = nullptr;
    dcph_model *M = new dcph_model();
    try {
        std::string wd = workdir ? workdir : ".";
        if (!wd.empty() && wd.back() != '/') wd += '/';
        // ---- PART 1: project.dat (dyna_project.f90:82-121) ----
        Unit u10(wd + "project.dat");
        u10.skip_record();
        u10.begin_list(); const int num_proj = u10.integer(); u10.end_list();
        if (id_project < 1 || id_project > num_proj) die("project id %d out of range (1..%d)", id_project, num_proj);
        std::string proj_dir, proj_file, proj_out;
        for (int i = 1; i <= num_proj; ++i) {
            u10.skip_record();
            const std::string name = field_a700(u10.read_record());
            const std::string dir = field_a700(u10.read_record());
            const std::string file = field_a700(u10.read_record());
            const std::string outd = field_a700(u10.read_record());
            (void)name;
            if (i == id_project) { proj_dir = dir; proj_file = file; proj_out = outd; }
        }
        // a relative project_dir is taken relative to the directory of project.dat
        if (proj_dir.empty() || proj_dir[0] != '/') proj_dir = wd + proj_dir;
        M->proj_dir = proj_dir;
        M->log_path = proj_dir + "/OUTPUT/DYMOND.log";                    // :136
        // ---- PART 2: settings file (:154-234) ----
        const std::string settings = proj_dir + proj_file + ".dat";
        Unit us(settings);
        us.skip_record(); us.skip_record();
        const std::string mc_par_file = field_a700(us.read_record());
        const std::string mstruct_file = field_a700(us.read_record());
        const int num_cat = field_i4(us.read_record(), settings);
        const int num_flow = field_i4(us.read_record(), settings);
        us.skip_record(); us.skip_record();
        if (id_cat < 1 || id_cat > num_cat) die("catchment setup id %d out of range (1..%d)", id_cat, num_cat);
        if (id_flow < 1 || id_flow > num_flow) die("input setup id %d out of range (1..%d)", id_flow, num_flow);
        std::string cat_file, meta_file, flow_conn, riv_data, flow_point;
        for (int i = 1; i <= num_cat; ++i) {
            us.skip_record();
            us.read_record();                                             // setup_name
            const std::string a = field_a700(us.read_record()), b = field_a700(us.read_record()),
                              c = field_a700(us.read_record()), d = field_a700(us.read_record()), e = field_a700(us.read_record());
            if (i == id_cat) { cat_file = a; meta_file = b; flow_conn = c; riv_data = d; flow_point = e; }
        }
        us.skip_record(); us.skip_record();
        std::string flow_file, precip_file, evap_file;
        for (int i = 1; i <= num_flow; ++i) {
            us.skip_record();
            us.read_record();
            const std::string a = field_a700(us.read_record()), b = field_a700(us.read_record()), c = field_a700(us.read_record());
            if (i == id_flow) { flow_file = a; precip_file = b; evap_file = c; }   // flow, precip, evap in that order (:224-226)
        }
        M->out_prefix = proj_dir + proj_out + "_";                        // :322-323
        // ---- the reference's setup sequence (dyna_main.f90:95-160) ----
        read_tread_dyna(*M, proj_dir + meta_file, proj_dir + cat_file);
        read_inputs(*M, proj_dir + flow_file, proj_dir + precip_file, proj_dir + evap_file);
        read_mc_setup(*M, proj_dir + mc_par_file);
        read_modelstruct(*M, proj_dir + mstruct_file);
        read_routingdata(*M, proj_dir + flow_conn, proj_dir + riv_data, proj_dir + flow_point);
        finish_desc(*M);
        *out = M;
        return DCP_OK;
    } catch (const std::exception &e) {
        g_err = e.what();
        delete M;
        return DCP_ERR_INVALID;
    }
}

// ------------------------------------------------------------------------------------------------
// Standalone router front end (DTA/route_run_standalone.f90:83-189): route_processing_read_info (flow tree + river data) and the
// ASCII-grid flow series (dta_utility.f90:519-580: one ROW per timestep, one COLUMN per input point).  The series has one column
// per node, or one column less when the first node is an ungauged sea outlet (route_run_standalone.f90:139-151: column i then
// feeds node i + 1).  The HRU part of the descriptor is a placeholder (one HRU per input, no forcing): dcp_route_timeseries
// only uses the river network.
// ------------------------------------------------------------------------------------------------
int dcph_load_router(const char *river_data_file, const char *flow_conn_file, const char *timeseries_file, dcph_model **out) {
    if (!river_data_file || !flow_conn_file || !timeseries_file || !out) { g_err = "null argument"; return DCP_ERR_INVALID; }
    *out = nullptr;
    try {
        std::unique_ptr<dcph_model> M(new dcph_model());
        const std::string conn = flow_conn_file;
        if (conn.size() < 14) die("-tree: '%s' is not a *_flow_conn.txt file", conn.c_str());
        const std::string point = conn.substr(0, conn.size() - 14) + "_flow_point.txt";          // dta_route_processing.f90:72
        // ---- ASCII grid (read_ascii_grid): header of 5 or 6 records, then nrows records of ncols values ----
        Unit u(timeseries_file);
        auto header = [&](const char *name) {
            u.begin_list();
            const std::string key = u.token();
            std::string a = key, b = name;
            for (auto &c : a) c = (char)tolower(c);
            for (auto &c : b) c = (char)tolower(c);
            if (a != b) die("invalid ascii grid '%s': expected %s, found %s", timeseries_file, name, key.c_str());
            const double v = u.real();
            u.end_list();
            return v;
        };
        const int ncols = (int)header("ncols"), nrows = (int)header("nrows");
        header("xllcorner"); header("yllcorner"); header("cellsize");
        {   // optional nodata record (dta_utility.f90:548-556): anything else is already grid data
            const std::string rec = u.peek_record();
            std::string low = rec;
            for (auto &c : low) c = (char)tolower(c);
            if (low.find("nodata") != std::string::npos) u.skip_record();
        }
        if (ncols < 1 || nrows < 1) die("ascii grid '%s': ncols / nrows must be positive", timeseries_file);
        M->router_inflow.assign((size_t)ncols * nrows, 0.0);
        for (int t = 0; t < nrows; ++t) {
            u.begin_list();
            for (int i = 0; i < ncols; ++i) M->router_inflow[(size_t)t * ncols + i] = u.real();     // ABI layout: input column fastest
            u.end_list();
        }
        // ---- placeholder HRU part: one HRU per input column ----
        dcp_catchment_desc &d = M->d;
        const int nriv = ncols;
        d.nac = nriv; d.nriv = nriv; d.npt = 1; d.nstep = nrows; d.ngp = 1; d.nge = 1; d.ntt = 1; d.dt = 1.0;
        M->ac.assign(nriv, 1.0 / nriv); M->st.assign(nriv, 8.0); M->mslp.assign(nriv, 0.1);
        M->ippt.assign(nriv, 1); M->ipet.assign(nriv, 1); M->ipar.assign(nriv, 1); M->ipriv.resize(nriv);
        M->ims_rz.assign(nriv, 1); M->ims_uz.assign(nriv, 1); M->ntrans.assign(nriv, 0); M->itrans.clear();
        M->wtrans.assign(nriv, 1.0); M->rivers.assign((size_t)nriv * nriv, 0.0);
        for (int i = 0; i < nriv; ++i) { M->ipriv[i] = i + 1; M->rivers[(size_t)i * nriv + i] = 1.0; }
        M->sum_ac_riv.assign(nriv, 1.0 / nriv); M->qobs_start.assign(nriv, 1e-5);
        M->rain.assign(nrows, 0.0); M->pet.assign(nrows, 0.0);
        read_routingdata(*M, conn, river_data_file, point);
        const int nnode = d.nnode;
        if (nriv == nnode) {                                                   // "flow file points matches nodes"
            for (int i = 0; i < nriv; ++i) M->node_to_flow_mapping[i] = i + 1;
        } else if (nriv + 1 == nnode) {                                        // no column for the sea outlet: column i -> node i + 1
            for (int i = 0; i < nriv; ++i) M->node_to_flow_mapping[i] = i + 2;
        } else {
            die("timeseries_flow_input does not match the _flow_conn.txt: %d columns for %d nodes "
                "(timeseries data must have one column for each node)", nriv, nnode);
        }
        for (int i = 0; i < nnode; ++i) M->frac_sub_area[i] = 1.0;
        // output headers: gauge ids of the gauge nodes, in node order (route_processing_init, dta_route_processing.f90:141-148)
        for (int i = 0; i < nnode; ++i)
            if (M->node_type[i] == 2) M->router_headers.push_back(std::to_string(M->node_gauge_id[i]));
        if ((int)M->router_headers.size() != nriv)
            die("route_process_run_timeseries: error size mismatch (%d gauge nodes, %d flow columns)", (int)M->router_headers.size(), nriv);
        finish_desc(*M);
        d.qobs = nullptr; d.nstep_obs = 0;
        *out = M.release();
        return DCP_OK;
    } catch (const std::exception &e) {
        g_err = e.what();
        return DCP_ERR_INVALID;
    }
}

const double *dcph_router_inflow(const dcph_model *m) { return m ? m->router_inflow.data() : nullptr; }

// write_numeric_list (DTA/dta_utility.f90:402-470): tab-separated headers, then one record per timestep in
// '(1(F0.p) n-1(1x,F0.p))' -- gfortran's F0.d drops the leading zero
int dcph_write_routed(const dcph_model *m, const double *routed /*[nriv x nstep]*/, const char *path, int precision, int f0_leading_zero) {
    if (!m || !routed || !path) { g_err = "null argument"; return DCP_ERR_INVALID; }
    FILE *f = fopen(path, "w");
    if (!f) { g_err = std::string("error opening output file: ") + path; return DCP_ERR_INVALID; }
    std::string hdr;
    for (size_t i = 0; i < m->router_headers.size(); ++i) hdr += (i ? "\t" : "") + m->router_headers[i];
    fprintf(f, "%s\n", hdr.c_str());
    const int nriv = m->d.nriv;
    for (int t = 0; t < m->d.nstep; ++t) {
        std::string rec;
        for (int i = 0; i < nriv; ++i) rec += (i ? " " : "") + fmt_F0(routed[(size_t)t * nriv + i], precision, f0_leading_zero != 0);
        fprintf(f, "%s\n", rec.c_str());
    }
    fclose(f);
    return DCP_OK;
}

// "-write-members 1:10,250,9000:9003" (decipher_b200_run): 1-based member ids, ranges inclusive -> ascending unique 0-based indexes.
// Returns the count (<= cap entries are written), or -1 with dcph_last_error() set.
int64_t dcph_parse_member_list(const char *list, int64_t numsim, int64_t *out, int64_t cap) {
    if (!list) { g_err = "null member list"; return -1; }
    std::vector<int64_t> sel;
    const char *p = list;
    while (*p) {
        char *e = nullptr;
        long long a = strtoll(p, &e, 10), b = a;
        if (e == p) { g_err = std::string("-write-members: cannot parse '") + p + "'"; return -1; }
        if (*e == ':') {
            p = e + 1;
            b = strtoll(p, &e, 10);
            if (e == p) { g_err = "-write-members: range without an end"; return -1; }
        }
        if (a < 1 || b < a || b > numsim) {
            g_err = "-write-members: " + std::to_string(a) + ":" + std::to_string(b) + " outside 1.." + std::to_string(numsim);
            return -1;
        }
        for (long long m = a; m <= b; ++m) sel.push_back(m - 1);
        p = e;
        if (*p == ',') ++p;
        else if (*p) { g_err = std::string("-write-members: unexpected '") + *p + "'"; return -1; }
    }
    std::sort(sel.begin(), sel.end());
    sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
    for (int64_t i = 0; i < (int64_t)sel.size() && i < cap; ++i) out[i] = sel[i];
    return (int64_t)sel.size();
}

void dcph_free(dcph_model *m) { delete m; }
const dcp_catchment_desc *dcph_desc(const dcph_model *m) { return m ? &m->d : nullptr; }
int dcph_numsim(const dcph_model *m) { return m ? m->numsim : 0; }
void dcph_seeds(const dcph_model *m, int *s1, int *s2) { if (m) { *s1 = m->seed_1; *s2 = m->seed_2; } }
const double *dcph_bounds_ll(const dcph_model *m) { return m ? m->ll.data() : nullptr; }
const double *dcph_bounds_ul(const dcph_model *m) { return m ? m->ul.data() : nullptr; }
const char *dcph_out_prefix(const dcph_model *m) { return m ? m->out_prefix.c_str() : ""; }
const char *dcph_log_path(const dcph_model *m) { return m ? m->log_path.c_str() : ""; }

int dcph_genpar(const dcph_model *m, int64_t n_member, double *mcpar) {
    if (!m || !mcpar || n_member < 0) { g_err = "bad argument"; return DCP_ERR_INVALID; }
    Ranmar r;
    r.ranin(m->seed_1, m->seed_2);                                        // dyna_main.f90:139
    const int npt = m->d.npt;
    for (int64_t mm = 0; mm < n_member; ++mm) {
        double *p = mcpar + (size_t)mm * npt * 7;
        for (int ig = 0; ig < npt; ++ig)                                  // dyna_genpar.f90:35-41
            for (int igen = 0; igen < 7; ++igen) {
                const double u = r.next();
                const size_t o = (size_t)igen * npt + ig;
                p[o] = m->ll[o] + u * (m->ul[o] - m->ll[o]);
            }
    }
    return DCP_OK;
}

int dcph_write_member(const dcph_model *m, int64_t i_mc, const double *mcpar, const double *q, const double *totals,
                      const double *diag, int32_t status, int f0_leading_zero) {
    if (!m || !mcpar || !q || !totals || !diag) { g_err = "bad argument"; return DCP_ERR_INVALID; }
    char id[32];
    snprintf(id, sizeof(id), "mc_id_%08lld", (long long)i_mc);            // dyna_file_open.f90:26
    const std::string base = m->out_prefix + id;
    FILE *res = fopen((base + ".res").c_str(), "w");
    FILE *flow = res ? fopen((base + ".flow").c_str(), "w") : nullptr;
    if (!res || !flow) {
        if (res) fclose(res);
        g_err = "cannot open output files " + base + ".res/.flow (does the OUTPUT directory exist?)";
        return DCP_ERR_INVALID;
    }
    const int npt = m->d.npt, nriv = m->d.nriv, nstep = m->d.nstep;
    const bool lz = f0_leading_zero != 0;
    // init block (dyna_init_satzone.f90:177-191), list-directed records start with a blank
    fprintf(res, " \n ! INITIALISATION\n");
    const int sol = status & DCP_ST_SOLFLAG_MASK;
    if (sol == 1) fprintf(res, " Solution found!\n");
    else if (sol == 0) fprintf(res, " No solution found - going for the minimum solution\n");
    else fprintf(res, " Maximum QBF is less than mean_q_ac_weighted - deficits set to 0 and max QBFs\n");
    fprintf(res, " Starting qobs %s mm/day\n", fmt_list_real(diag[0] * 1000).c_str());
    fprintf(res, " Best estimate of QB riv%s mm/day\n", fmt_list_real(diag[1] * 1000).c_str());
    fprintf(res, " Finishing qt0 %s mm/day\n", fmt_list_real(diag[2] * 1000).c_str());
    fprintf(res, " \n");
    // parameters (dyna_write_output.f90:44-51)
    fprintf(res, "! PARAMETERS\n");
    const int aw[8] = {16, 10, 12, 10, 11, 12, 11, 10};
    std::string hdr = fmt_A("Param_class", aw[0]);
    for (int j = 0; j < 7; ++j) hdr += fmt_A(std::string(PM[j]) + "_def", aw[j + 1]);
    fprintf(res, "%s\n", hdr.c_str());
    const int fw[7] = {10, 12, 10, 10, 13, 11, 10};
    for (int i = 0; i < npt; ++i) {
        std::string row;
        char b[32];
        snprintf(b, sizeof(b), "%12d", i + 1);
        row = b;
        for (int j = 0; j < 7; ++j) row += fmt_F(mcpar[(size_t)j * npt + i], fw[j], 4);
        fprintf(res, "%s\n", row.c_str());
    }
    fprintf(res, "\n");
    // reach totals (:79-93)
    fprintf(res, "! SUMMARY STATS PER RIVER REACH\n\n");
    const char *titles[3] = {"! PRECIP TOTALS (mm)", "! PET TOTALS (mm)", "! ET TOTALS (mm)"};
    for (int k = 0; k < 3; ++k) {
        fprintf(res, "%s\n", titles[k]);
        std::string row;
        for (int r = 0; r < nriv; ++r) row += " " + fmt_F0(totals[k + 3 * r], 2, lz);
        fprintf(res, "%s\n", row.c_str());
        if (k < 2) fprintf(res, "\n");
    }
    fclose(res);
    // flows (:96-100): nriv(1X,f0.8) per step
    std::string line;
    for (int t = 0; t < nstep; ++t) {
        line.clear();
        for (int r = 0; r < nriv; ++r) line += " " + fmt_F0(q[(size_t)t * nriv + r], 8, lz);
        fputs(line.c_str(), flow);
        fputc('\n', flow);
    }
    fclose(flow);
    return DCP_OK;
}

}  // extern "C"
