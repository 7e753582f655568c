"""Writers for the reference's on-disk text formats, used to emit synthetic projects
that `program main` — and our C++ host (csrc/host/dcp_host.cpp) — can read:

  project.dat, settings.dat      RRMODEL/dyna_project.f90:82-229 (format(a13,1X,a700) / (a13,1X,i4))
  params.dat                     RRMODEL/dyna_mc_setup.f90:81-110
  model_structure.dat            RRMODEL/dyna_modelstruct_setup.f90:30-54
  *_hru_meta.dat, *_dyna_hru.dat DTA/dta_calculate_hrus.f90:327-354, 530-640 (readers dyna_tread_dyna.f90:46-170)
  *_flow_point/_flow_conn/_riv_data.dat   DTA/dta_preprocess.f90:279-341
  rain / PET / observed flow     RRMODEL/dyna_inputs.f90:54-160
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

PARAM_NAMES = ["szm", "lnto", "srmax", "srinit", "chv", "td", "smax"]


def _lab(label: str, value: str) -> str:
    assert len(label) == 13, label
    return f"{label}\t{value}\n"


def write_project(raw: Dict, root: str, ll: np.ndarray, ul: np.ndarray, numsim: int = 10, seeds=(5, 2000),
                  name: str = "Catch1") -> str:
    """Writes a complete single-catchment project under `root` and returns `root` (the directory that
    holds project.dat).  project_dir is written relative ('./'), resolved against that directory."""
    os.makedirs(os.path.join(root, "SETTINGS"), exist_ok=True)
    os.makedirs(os.path.join(root, "INPUTS"), exist_ok=True)
    os.makedirs(os.path.join(root, "OUTPUT"), exist_ok=True)
    nac, nriv, npt, nstep = raw["nac"], raw["nriv"], raw["npt"], raw["nstep"]
    with open(os.path.join(root, "project.dat"), "w") as f:
        f.write("! DYMOND synthetic project file\n1\t\t! total number of project files\n\n")
        f.write(_lab("project_name:", name) + _lab("project_dir :", "./") + _lab("projectfiles:", "SETTINGS/settings")
                + _lab("project_out :", f"OUTPUT/{name}"))
    with open(os.path.join(root, "SETTINGS", "settings.dat"), "w") as f:
        f.write("! DYMOND Settings File\n\n")
        f.write(_lab("param_file  :", "SETTINGS/params.dat") + _lab("mstruct_file:", "SETTINGS/model_structure.dat"))
        f.write(_lab("n_cat_setups:", "0001") + _lab("n_flow_setps:", "0001"))
        f.write("\n --- DATA LOCATIONS AND SETTINGS FOR ALL CATCHMENT SETUPS ---\n\n")
        f.write(_lab("setup_name  :", "synthetic HRU setup") + _lab("HRU_filepath:", f"INPUTS/{name}_dyna_hru.dat")
                + _lab("HRU_metafile:", f"INPUTS/{name}_hru_meta.dat") + _lab("flow_conn   :", f"INPUTS/{name}_flow_conn.dat")
                + _lab("riv_data    :", f"INPUTS/{name}_riv_data.dat") + _lab("flow_point  :", f"INPUTS/{name}_flow_point.dat"))
        f.write("\n --- DATA LOCATION AND SETTINGS FOR ALL INPUT SETUPS ---\n\n")
        f.write(_lab("setup_name  :", "synthetic forcing") + _lab("flow_file   :", f"INPUTS/{name}_obsq.dat")
                + _lab("precip_file :", f"INPUTS/{name}_rain.dat") + _lab("evap_file   :", f"INPUTS/{name}_PET.dat"))
    with open(os.path.join(root, "SETTINGS", "params.dat"), "w") as f:
        f.write(f"{numsim}\t\t! Number of simulations you want to run\n{seeds[0]}\t{seeds[1]}\t! Seeds for this run\n")
        f.write(f"{raw['ntt']}\t0.8\t0.000000001\t! Kinematic solution Parameters\n")
        f.write(f"{npt}\t14\t!number of parameter layers, number of param names\n\n")
        f.write("param_layer\t" + "\t".join(f"{p}_ll\t{p}_ul" for p in PARAM_NAMES) + "\n")
        for i in range(npt):
            f.write(f"{i + 1}\t" + "\t".join(f"{float(ll[i, j])!r}\t{float(ul[i, j])!r}" for j in range(7)) + "\n")
    with open(os.path.join(root, "SETTINGS", "model_structure.dat"), "w") as f:
        f.write("! MODEL STRUCTURES\n \n1\t2\t! number of model structure types, number of components\n")
        f.write("model_layer\trootzone\tunsatzone\n1\t1\t1\n")
    inp = os.path.join(root, "INPUTS")
    with open(os.path.join(inp, f"{name}_hru_meta.dat"), "w") as f:
        f.write("! HRU meta file (synthetic)\n")
        f.write(f"HRUs\t{nac}\nNUM_COLS\t9\n!\n")
        f.write("HRU_ID\tNUM_CELLS\tMEAN_ATB\tMEAN_SLOPE\tGAUGE_ID\tRAIN_ID\tPET_ID\tPARAM_ID\tMODELSTRUCT_ID\n")
        for i in range(nac):
            f.write(f"{i + 1}\t{int(raw['cells'][i])}\t{raw['st'][i]:.5f}\t{raw['mslp'][i]:.8f}\t{int(raw['ipriv'][i])}\t"
                    f"{int(raw['ippt'][i])}\t{int(raw['ipet'][i])}\t{int(raw['ipar'][i])}\t{int(raw['ims'][i])}\n")
    with open(os.path.join(inp, f"{name}_dyna_hru.dat"), "w") as f:
        f.write("! HRU flux file (synthetic)\n")
        f.write(f"HRUs\t{nac}\nCatchment_Area\t{float(raw['cells'].sum()) * 2.5e-3:.4f}\nClassifying_Layers\t2\nGauges\t{nriv}\n")
        f.write("ATB_Classes\t10\nHRU FLUXES\n")               # L = 2 non-blank records: L-1 class lines + the banner
        off_t = off_w = 0
        for i in range(nac):
            nt = int(raw["ntrans"][i])
            f.write(f"HRU\t{i + 1}\tSharing to\t{nt}\tHRUs\n")
            cum = 0.0
            for k in range(nt):
                w = raw["wtrans"][off_w + k]
                cum += w
                f.write(f"{int(raw['itrans'][off_t + k])}\t{w:.6f}\t{cum:.6f}\n")
            w = raw["wtrans"][off_w + nt]
            f.write(f"RIVS\t{w:.6f}\t{cum + w:.6f}\n")
            off_t += nt
            off_w += nt + 1
        f.write("HRU RIV FLUXES\n")
        for i in range(nac):
            rid, share = raw["riv_share"][i]
            f.write(f"HRU\t{i + 1}\tSharing to\t1\tRIVS\n{rid}\t{share:.6f}\t{share:.6f}\n")
    np.savetxt(os.path.join(inp, f"{name}_flow_point.dat"), raw["flow_point"], fmt="%.3f", delimiter="\t",
               header="Node_ID\tx_easting\ty_northing\tType\tdist\th", comments="")
    np.savetxt(os.path.join(inp, f"{name}_flow_conn.dat"), raw["flow_conn"], fmt="%.3f", delimiter="\t",
               header="Node_ID\tx\ty\tType\tID_down\tx2\ty2\tdelta_dist\tdelta_h\tslope", comments="")
    np.savetxt(os.path.join(inp, f"{name}_riv_data.dat"), raw["river_data"], fmt="%.3f", delimiter="\t",
               header="riv_id\tarea\tdist\tsection_dist\televation\tslope\tds_dist\tds_elevation", comments="")

    def series(path, data, dec, flow_err=None):
        n = data.shape[0]
        with open(path, "w") as f:
            f.write("! synthetic series\n! units: m per timestep\n!\n")
            f.write(f"{float(raw['dt'])!r}\t{data.shape[1]}\t{n}" + (f"\t{float(flow_err)!r}" if flow_err is not None else "") + "\n")
            f.write("!\nY\tM\tD\tH\t" + "\t".join(f"id{k + 1}" for k in range(n)) + "\n")
            hours = np.arange(data.shape[1])
            for t in range(data.shape[1]):
                d, h = divmod(int(hours[t]), 24)
                f.write(f"2001\t1\t{d + 1}\t{h}\t" + "\t".join(f"{v:.{dec}f}" for v in data[:, t]) + "\n")
    series(os.path.join(inp, f"{name}_rain.dat"), raw["rain"], 8)
    series(os.path.join(inp, f"{name}_PET.dat"), raw["pet"], 9)
    series(os.path.join(inp, f"{name}_obsq.dat"), raw["qobs"], 9, raw["flow_err"])
    with open(os.path.join(root, "auto.dat"), "w") as f:
        f.write("1\n1\n1\n")
    return root


def read_perf_table(path: str) -> Dict:
    """<prefix>ensemble_perf.bin written by decipher_b200_run (v2): the sampled parameter sets (after the srinit clip of
    dyna_init_satzone.f90:203-206) next to their performance measures and status words."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"DCPPERF2", raw[:8]
    numsim, nriv, nmeasure, npt = np.frombuffer(raw[8:40], dtype=np.int64).tolist()
    o = 40
    n = npt * 7 * numsim
    mcpar = np.frombuffer(raw[o:o + 8 * n], dtype=np.float64).reshape((npt, 7, numsim), order="F"); o += 8 * n
    n = nmeasure * nriv * numsim
    perf = np.frombuffer(raw[o:o + 8 * n], dtype=np.float64).reshape((nmeasure, nriv, numsim), order="F"); o += 8 * n
    status = np.frombuffer(raw[o:o + 4 * numsim], dtype=np.int32)
    return dict(mcpar=mcpar, perf=perf, status=status)


def read_flow_table(path: str) -> Dict:
    """<prefix>ensemble_flow.bin (decipher_b200_run -flow-bin): FP64 flow series q(nriv, nstep) of the selected members."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"DCPFLOW1", raw[:8]
    nsel, nriv, nstep = np.frombuffer(raw[8:32], dtype=np.int64).tolist()
    ids = np.frombuffer(raw[32:32 + 8 * nsel], dtype=np.int64)
    q = np.frombuffer(raw[32 + 8 * nsel:], dtype=np.float64).reshape((nriv, nstep, nsel), order="F")
    return dict(members=ids, q=q)


def write_router_inputs(cat, inflow: np.ndarray, root: str, stem: str = "net") -> Dict[str, str]:
    """Files of the standalone router (DTA/route_run_standalone.f90): <stem>_flow_conn.txt + <stem>_flow_point.txt (routing tree),
    <stem>_river_data.txt and the flow series as an ASCII grid (one row per timestep, one column per input point).
    `cat` supplies the network (node ids / types, meta['down'], river_data); inflow is (n_points, nstep)."""
    os.makedirs(root, exist_ok=True)
    down = np.asarray(cat.meta["down"])
    gid = np.asarray(cat.node_gauge_id)
    nn = len(gid)
    fp = np.array([(gid[n], 1000.0 + n, 2000.0 + n, cat.node_type[n], 0.0, 10.0) for n in range(nn)], dtype=np.float64)
    fc = np.array([(gid[n], 1000.0 + n, 2000.0 + n, cat.node_type[n], gid[down[n] - 1] if down[n] > 0 else 0, 0.0, 0.0, 0.0, 1.0, 0.002)
                   for n in range(nn)], dtype=np.float64)
    paths = {k: os.path.join(root, f"{stem}_{k}.txt") for k in ("flow_conn", "flow_point", "river_data")}
    np.savetxt(paths["flow_point"], fp, fmt="%.3f", delimiter="\t", header="Node_ID\tx_easting\ty_northing\tType\tdist\th", comments="")
    np.savetxt(paths["flow_conn"], fc, fmt="%.3f", delimiter="\t",
               header="Node_ID\tx\ty\tType\tID_down\tx2\ty2\tdelta_dist\tdelta_h\tslope", comments="")
    np.savetxt(paths["river_data"], cat.river_data, fmt="%.3f", delimiter="\t",
               header="riv_id\tarea\tdist\tsection_dist\televation\tslope\tds_dist\tds_elevation", comments="")
    paths["timeseries"] = os.path.join(root, f"{stem}_flow.txt")
    n_points, nstep = inflow.shape
    with open(paths["timeseries"], "w") as f:
        f.write(f"ncols {n_points}\nnrows {nstep}\nxllcorner 0\nyllcorner 0\ncellsize 1\nNODATA_value -9999\n")
        for t in range(nstep):
            f.write(" ".join(repr(float(v)) for v in inflow[:, t]) + "\n")
    return paths
