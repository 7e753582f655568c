"""The reference's own shipped control files (RRMODEL_EXAMPLE_FILES/) through the C++ host reader, VERBATIM.

These four files are the only fixtures the reference holds for the Monte-Carlo path.
params.dat, model_structure.dat and settings.dat are staged byte for byte; project.dat only has its two
`project_dir` values (absolute paths of the authors' machine, /data/MODEL/CatchmentN/) pointed at the
staging directories.  The data files settings.dat names (INPUTS/Catch1_*) are not shipped upstream, so a
synthetic catchment with three parameter layers is written under exactly those names.

Checked against the text of the files: numsim = 10, seeds 5 2000, NTT = 1 (dyna_mc_setup.f90:81-110),
the 3 x 14 bounds, the model-structure table (dyna_modelstruct_setup.f90:30-111), the 3 catchment and
3 input setups of settings.dat (dyna_project.f90:170-229) and the two projects of project.dat (:82-125).
"""
import filecmp
import os
import sys

import numpy as np
import pytest

from decipher_b200 import formats, host, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import f90_cases  # noqa: E402

GOLD = os.path.join(HERE, "golden", "ref_example")
REF = "/root/reference/RRMODEL_EXAMPLE_FILES"
FILES = ["params.dat", "model_structure.dat", "settings.dat", "project.dat"]

# params.dat:7-9 as printed in the file (ll, ul per parameter: szm lnto srmax srinit chv td smax)
BOUNDS = np.array([
    [0.001, 0.07, -7, 5, 0.005, 0.15, 0, 0.01, 250, 4000, 0.1, 40, 0.2, 3],
    [0.005, 0.007, -7, -2, 0.005, 0.15, 0, 0.01, 1000, 4000, 0.1, 1, 0.2, 3],
    [0.01, 0.04, 1, 4, 0.005, 0.15, 0, 0.01, 3000, 4000, 0.1, 40, 0.2, 3]])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_golden_copies_are_the_reference_files():
    for f in FILES:
        assert filecmp.cmp(os.path.join(GOLD, f), os.path.join(REF, f), shallow=False), f


@pytest.fixture(scope="module")
def staged(tmp_path_factory):
    """work/project.dat (reference text, project_dir substituted) + Catchment1/, Catchment2/ holding the verbatim
    SETTINGS files and synthetic INPUTS under the names settings.dat lists."""
    work = str(tmp_path_factory.mktemp("refproj"))
    dirs, raws = f90_cases.stage_reference_example(work, GOLD)
    return work, dirs, raws


def test_reference_params_and_model_structure_verbatim(staged):
    work, dirs, raws = staged
    p = host.Project(work, 1, 1, 1)
    assert p.numsim == 10                                   # params.dat:1
    assert p.seeds == (5, 2000)                             # params.dat:2
    assert p.desc.ntt == 1                                  # params.dat:3, first field (NTT)
    assert p.desc.npt == 3                                  # params.dat:4
    assert p.ll.shape == (3, 7) and p.ul.shape == (3, 7)
    assert np.array_equal(p.ll, BOUNDS[:, 0::2]) and np.array_equal(p.ul, BOUNDS[:, 1::2])      # params.dat:7-9, 3 x 14
    ll, ul = synth.example_bounds(3)                        # the bounds every synthetic benchmark draws from
    assert np.array_equal(p.ll, ll) and np.array_equal(p.ul, ul)
    cat = p.catchment()
    assert (cat.ims_rz == 1).all() and (cat.ims_uz == 1).all()          # model_structure.dat:5  "1 1 1"
    assert p.out_prefix == dirs[0] + "OUTPUT/Catch1_" + "_"             # dyna_project.f90:322-323
    # the sampled parameter table of the shipped configuration: first draw of RANMAR(5, 2000) on layer 1
    mc = p.genpar()
    assert mc.shape == (3, 7, 10)
    u = host.ranmar(5, 2000, 0, 21)
    assert mc[0, 0, 0] == 0.001 + u[0] * (0.07 - 0.001)                  # dyna_genpar.f90:35-41: ll + u*(ul-ll)
    assert np.array_equal(mc, synth.sample_parameters(3, 10))
    p.close()


@pytest.mark.parametrize("ids", [(1, 1, 1), (1, 2, 2), (1, 3, 3), (2, 1, 1), (2, 3, 2)])
def test_reference_settings_and_project_selection(staged, ids):
    """settings.dat's 3 catchment setups x 3 input setups and project.dat's 2 projects select the right files."""
    work, dirs, raws = staged
    ip, ic, iflow = ids
    p = host.Project(work, ip, ic, iflow)
    want_hru = synth.assemble(raws[(ip - 1, ic - 1)])
    want_forc = synth.assemble(raws[(ip - 1, iflow - 1)])
    got = p.catchment()
    assert got.nac == want_hru.nac and np.array_equal(got.ac, want_hru.ac) and np.array_equal(got.st, want_hru.st)
    assert np.array_equal(got.wtrans, want_hru.wtrans) and np.array_equal(got.itrans, want_hru.itrans)
    assert got.nstep == want_forc.nstep and np.array_equal(got.rain, want_forc.rain) and np.array_equal(got.pet, want_forc.pet)
    assert p.out_prefix.startswith(dirs[ip - 1] + "OUTPUT/Catch%d_" % ip)
    p.close()


def test_reference_settings_rejects_out_of_range_setup(staged):
    work, *_ = staged
    with pytest.raises(Exception):
        host.Project(work, 1, 4, 1)                 # n_cat_setups = 0003
    with pytest.raises(Exception):
        host.Project(work, 3, 1, 1)                 # 2 projects
