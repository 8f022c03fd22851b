"""The drop-in itself: the reference's OWN library and binary -- XML input, data
loading, source sampling, synchronize_bank, k statistics, tally accumulation, statepoint
writing, all unmodified -- with the three hook edits of openmc_b200/host/apply_hooks.py and the
host side of the B200 path (openmc_b200/host/omcgpu_host.cpp) compiled in
(oracle/_ref/libopenmc_gpu.so, oracle/_ref/openmc_gpu; recipe oracle/Makefile).

* the C++ flattener (model:: / data:: -> omcgpu_model) gives, field by field, the arrays this
  repo's Python flattener derives from the same description;
* `OPENMC_GPU=1 openmc_gpu` writes the statepoint `openmc` writes: here, on the CPU tier, the
  device library is the TEST-ONLY host compile of the engine sources (tests/_emu, the same
  event loop and device functions with glibc's libm), so k of every generation and the source
  bank are bit-identical and tallies agree to summation order;
* with OPENMC_GPU unset the patched binary is the reference;
* a model the device path refuses ends in the reference's fatal_error, not in a CPU fallback.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import cases
import reference_runs as rr
from openmc_b200 import _abi as abi
from openmc_b200 import mgxs, nuclide
from oracle import refrun
from oracle.binding import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BINARY = os.path.join(ROOT, "oracle", "_ref", "openmc_gpu")
GPU_LIB = os.path.join(ROOT, "oracle", "_ref", "libopenmc_gpu.so")


@pytest.fixture(scope="module")
def dropin(emu_lib):
    if os.path.isdir("/root/reference/src"):
        build()
    if not os.path.exists(GPU_BINARY):
        pytest.skip("oracle/_ref/openmc_gpu not built (no /root/reference on this box)")
    return emu_lib


def _inputs(name):
    m, lib, model, box, of, ce = rr.build(name)
    particles, batches, inactive = rr.RUNS[name]
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    rr.write_inputs(d, m, lib, box, of, ce, particles, batches, inactive)
    return d


def _run(d, engine_lib, gpu=True, threads=1):
    env = dict(os.environ)
    env.pop("OPENMC_GPU", None)
    if gpu:
        env["OPENMC_GPU"] = "1"
        env["OMCGPU_LIB"] = engine_lib
    for f in os.listdir(d):
        if f.startswith("statepoint"):
            os.remove(os.path.join(d, f))
    res = subprocess.run([GPU_BINARY, "-s", str(threads)], cwd=d, capture_output=True, text=True, env=env, timeout=1200)
    return res


@pytest.mark.parametrize("name", list(rr.RUNS))
def test_dropin_binary_equals_the_reference_binary(dropin, name):
    d = _inputs(name)
    try:
        res = _run(d, dropin)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        sp = [f for f in os.listdir(d) if f.startswith("statepoint")]
        from oracle import omh5
        got = rr.digest({"statepoint": omh5.File(os.path.join(d, sorted(sp)[-1]))})
    finally:
        shutil.rmtree(d, ignore_errors=True)
    ref = dict(np.load(rr.GOLDEN % name))     # what the unmodified binary writes (checked current elsewhere)
    assert np.array_equal(got["k_generation"], ref["k_generation"])
    assert np.array_equal(got["k_combined"], ref["k_combined"])
    assert np.array_equal(got["global_tallies"], ref["global_tallies"])
    for k in ("bank_r", "bank_u", "bank_E", "bank_wgt", "bank_delayed_group"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["tally_sum"].size == ref["tally_sum"].size
    if ref["tally_sum"].size:
        np.testing.assert_allclose(got["tally_sum"], ref["tally_sum"], rtol=1e-11, atol=1e-14 * ref["tally_sum"].max())
        np.testing.assert_allclose(got["tally_sum_sq"], ref["tally_sum_sq"], rtol=1e-11,
                                   atol=1e-14 * ref["tally_sum_sq"].max())


def test_several_generations_per_batch(dropin):
    """<generations_per_batch>: the seeds of a generation come from total_gen + overall_generation()
    (src/simulation.cpp:613-616, src/eigenvalue.cpp:152-178) and the tallies of a batch accumulate
    over its generations -- the drop-in's statepoint equals the reference binary's"""
    from oracle import omh5
    m, lib, model, box, of, ce = rr.build("pincell_mg")
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    try:
        rr.write_inputs(d, m, lib, box, of, ce, 400, 4, 1, gen_per_batch=3)
        out = {}
        for gpu in (True, False):
            res = _run(d, dropin, gpu=gpu)
            assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
            sp = sorted(f for f in os.listdir(d) if f.startswith("statepoint"))[-1]
            out[gpu] = rr.digest({"statepoint": omh5.File(os.path.join(d, sp))})
        a, b = out[True], out[False]
        assert a["k_generation"].size == 12 and np.array_equal(a["k_generation"], b["k_generation"])
        assert np.array_equal(a["k_combined"], b["k_combined"])
        for k in ("bank_r", "bank_E"):
            assert np.array_equal(a[k], b[k]), k
        np.testing.assert_allclose(a["tally_sum"], b["tally_sum"], rtol=1e-11, atol=1e-14 * b["tally_sum"].max())
        np.testing.assert_allclose(a["tally_sum_sq"], b["tally_sum_sq"], rtol=1e-11, atol=1e-14 * b["tally_sum_sq"].max())
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_restart_from_a_statepoint(dropin):
    """`openmc -r statepoint.3.h5`: the reference reloads k, the tallies and the source bank and goes on
    with batch 4 (load_state_point, src/state_point.cpp); the drop-in continues from the same state
    -- seeds from simulation::total_gen, host source bank uploaded -- and ends where the
    uninterrupted reference run ends"""
    from oracle import omh5
    m, lib, model, box, of, ce = rr.build("pincell_mg")
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    try:
        rr.write_inputs(d, m, lib, box, of, ce, 400, 5, 2,
                        extra={"state_point": "<batches>3 5</batches>", "sourcepoint": "<batches>3 5</batches>"})
        assert _run(d, dropin, gpu=False).returncode == 0           # (removes no statepoint it needs: none yet)
        full = rr.digest({"statepoint": omh5.File(os.path.join(d, "statepoint.5.h5"))})
        os.remove(os.path.join(d, "statepoint.5.h5"))
        env = dict(os.environ, OPENMC_GPU="1", OMCGPU_LIB=dropin)
        res = subprocess.run([GPU_BINARY, "-s", "1", "-r", "statepoint.3.h5"], cwd=d, capture_output=True, text=True,
                             env=env, timeout=1200)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        got = rr.digest({"statepoint": omh5.File(os.path.join(d, "statepoint.5.h5"))})
        assert np.array_equal(got["k_generation"], full["k_generation"])
        assert np.array_equal(got["bank_E"], full["bank_E"]) and np.array_equal(got["bank_r"], full["bank_r"])
        np.testing.assert_allclose(got["k_combined"], full["k_combined"], rtol=1e-12)
        np.testing.assert_allclose(got["tally_sum"], full["tally_sum"], rtol=1e-11, atol=1e-14 * full["tally_sum"].max())
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_switch_off_is_the_reference(dropin):
    d = _inputs("pincell_mg")
    try:
        res = _run(d, dropin, gpu=False)
        assert res.returncode == 0, res.stderr[-1500:]
        from oracle import omh5
        sp = sorted(f for f in os.listdir(d) if f.startswith("statepoint"))[-1]
        got = rr.digest({"statepoint": omh5.File(os.path.join(d, sp))})
    finally:
        shutil.rmtree(d, ignore_errors=True)
    ref = dict(np.load(rr.GOLDEN % "pincell_mg"))
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k


def test_refused_model_is_a_fatal_error(dropin):
    """a boundary with an albedo: the device path refuses it, the run ends with the reference's
    ' ERROR: ' convention (src/error.cpp:113-140) instead of silently using the CPU"""
    m, lib, model, box, of, ce = rr.build("pincell_mg")
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    try:
        rr.write_inputs(d, m, lib, box, of, ce, 200, 3, 1)
        path = os.path.join(d, "model.xml")
        xml = open(path).read()
        assert 'boundary="reflective"' in xml
        open(path, "w").write(xml.replace('boundary="reflective"', 'boundary="reflective" albedo="0.9"', 1))
        res = _run(d, dropin)
        assert res.returncode != 0
        assert "ERROR" in res.stderr and "not supported by the device path" in res.stderr.replace("\n", " ").replace("  ", " ")
        assert _run(d, dropin, gpu=False).returncode == 0      # the reference itself runs the model
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_lost_particles_are_reported_like_the_reference(dropin):
    """a pin cell whose moderator stops short of the boundary: the drop-in prints one warning per
    lost particle (the device keeps the first 64 of a generation), counts them into
    simulation::n_lost_particles and ends the run by the reference's rule
    (src/particle.cpp:762-785) -- the reference itself aborts on the same model"""
    m, lib, model, box, of, ce = rr.build("pincell_mg")
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    try:
        rr.write_inputs(d, m, lib, box, of, ce, 200, 3, 1)
        path = os.path.join(d, "model.xml")
        xml = open(path).read()
        # the moderator cell loses its outer x boundary's neighbourhood: region ... -<x1> becomes ... -<xp>
        import re
        sid = max(int(i) for i in re.findall(r'<surface[^>]*\bid="(\d+)"', xml)) + 1
        assert '<surface ' in xml
        xml = xml.replace("<surface ", f'<surface id="{sid}" type="x-plane" coeffs="0.5"/>\n  <surface ', 1)
        cells = re.findall(r'<cell [^>]*>', xml)
        mod = cells[-1]
        reg = re.search(r'region="([^"]*)"', mod).group(1)
        xml = xml.replace(mod, mod.replace(f'region="{reg}"', f'region="{reg} -{sid}"'), 1)
        open(path, "w").write(xml)
        res = _run(d, dropin)
        out = res.stdout + res.stderr
        assert "was lost on the device" in out.replace("\n", " ").replace("  ", " "), out[-2000:]
        assert "Maximum number of lost particles has been reached" in out.replace("\n", " ").replace("  ", " ")
        assert res.returncode != 0
        ref = _run(d, dropin, gpu=False)
        assert ref.returncode != 0 and "Maximum number of lost particles" in (ref.stdout + ref.stderr).replace("\n", " ")
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("setting, word", [("track", "tracks"), ("surf_source_write", "surface source")])
def test_host_side_per_particle_features_are_refused(dropin, setting, word):
    """what the reference does on the host while it transports a particle (surface source banking,
    particle tracks, the overlap check ...) the device path does not do: the run stops naming the feature
    instead of leaving it out silently"""
    m, lib, model, box, of, ce = rr.build("pincell_mg")
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    try:
        value = "1 1 1" if setting == "track" else "<surface_ids>1</surface_ids><max_particles>100</max_particles>"
        rr.write_inputs(d, m, lib, box, of, ce, 200, 3, 1, extra={setting: value})
        res = _run(d, dropin)
        text = " ".join((res.stdout + res.stderr).split())
        assert res.returncode != 0 and "not supported by the device path" in text and word in text, text[-1500:]
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_properties_changed_through_the_c_api_reach_the_device(dropin):
    """a coupled calculation on the reference's C API (tests/coupled_worker.py): after two batches the
    fuel density is lowered and a cell heated through openmc_material_set_density /
    openmc_cell_set_temperature.  The glue compares the mutable arrays with what the device holds
    before every generation and sends the changes over (omcgpu_update): k-combined after every
    batch equals the reference library's doing the same -- and differs from an undisturbed run"""
    d = _inputs("pincell_ce")
    try:
        def run(gpu, change=True, worker="coupled_worker.py"):
            env = dict(os.environ)
            env.pop("OPENMC_GPU", None)
            env.pop("COUPLED_NO_CHANGE", None)
            if not change:
                env["COUPLED_NO_CHANGE"] = "1"
            if gpu:
                env.update(OPENMC_GPU="1", OMCGPU_LIB=dropin)
            res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", worker), GPU_LIB], cwd=d,
                                 capture_output=True, text=True, env=env, timeout=1200)
            assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
            line = [ln for ln in res.stdout.splitlines() if ln.startswith("KEFF")][0]
            return [float(x) for x in line.split()[1:]]
        k_gpu, k_ref, k_plain = run(True), run(False), run(True, change=False)
        assert len(k_gpu) == rr.RUNS["pincell_ce"][1]
        np.testing.assert_allclose(k_gpu, k_ref, rtol=1e-12)          # (sums in another order: last ulp)
        assert k_gpu[:2] == k_plain[:2] and abs(k_gpu[-1] - k_plain[-1]) > 1e-4   # the change is felt
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_shannon_entropy_of_the_device_fission_bank(dropin):
    """an <entropy_mesh>: the reference computes the Shannon entropy of simulation::fission_bank every
    generation (src/eigenvalue.cpp:579-607) and prints it beside k; in the drop-in that bank stays
    on the device, which counts its sites on the mesh (omcgpu_count_fission_sites) -- the lines of
    every generation (k AND entropy) must be the reference's, character for character"""
    import re
    m, lib, model, box, of, ce = rr.build("c5g7_core")
    import copy
    m2 = copy.deepcopy(m)
    m2.settings["entropy_mesh"] = ((6, 6), (box[0], box[1]), (box[3], box[4]))
    d = tempfile.mkdtemp(prefix="omc_dropin_")
    try:
        rr.write_inputs(d, m2, lib, box, of, ce, 600, 5, 2)
        gpu = _run(d, dropin)
        assert gpu.returncode == 0, gpu.stdout[-1500:] + gpu.stderr[-1500:]
        ref = _run(d, dropin, gpu=False)
        assert ref.returncode == 0
        lines = lambda out: [ln for ln in out.splitlines() if re.match(r"^\s+\d+/\d+\s", ln)]
        a, b = lines(gpu.stdout), lines(ref.stdout)
        assert len(a) == 5 and a == b
        assert all(len(ln.split()) >= 3 for ln in a)          # Bat./Gen., k, entropy
        assert 0.0 < float(a[0].split()[2]) < np.log2(36) + 1e-9
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_event_timers_are_fed_by_the_device(dropin):
    """print_runtime (src/output.cpp:439-446) lists the event timers of an event-based run: in the
    drop-in they carry the device's per-kernel times"""
    d = _inputs("pincell_mg")
    try:
        env = dict(os.environ, OPENMC_GPU="1", OMCGPU_LIB=dropin)
        res = subprocess.run([GPU_BINARY, "-s", "1", "-e"], cwd=d, capture_output=True, text=True, env=env, timeout=1200)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        for line in ("Surface crossings", "Collisions", "Particle death"):
            assert line in res.stdout
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------------------------------------
# C++ flattener == Python flattener, field by field
# ---------------------------------------------------------------------------------------------
def _region_truth(tokens, senses):
    """evaluate an infix region expression (surface tokens +-(index+1), OMCGPU_OP_*) for one
    assignment of surface senses"""
    LP, RP, CO, IN, UN = (abi.OMCGPU_OP_LEFT_PAREN, abi.OMCGPU_OP_RIGHT_PAREN, abi.OMCGPU_OP_COMPLEMENT,
                          abi.OMCGPU_OP_INTERSECTION, abi.OMCGPU_OP_UNION)
    pos = 0

    def primary():
        nonlocal pos
        t = int(tokens[pos]); pos += 1
        if t == CO:
            return not primary()
        if t == LP:
            v = union()
            pos += 1
            return v
        return senses[abs(t) - 1] == (t > 0)

    def inter():
        nonlocal pos
        v = primary()
        while pos < len(tokens) and (tokens[pos] == IN or tokens[pos] not in (RP, UN)):
            if tokens[pos] == IN:
                pos += 1
            v = primary() and v
        return v

    def union():
        nonlocal pos
        v = inter()
        while pos < len(tokens) and tokens[pos] == UN:
            pos += 1
            v = inter() or v
        return v
    return union() if len(tokens) else True


def _regions_equivalent(tok_a, ofs_a, tok_b, ofs_b, n_surfaces, seed=5):
    """the reference keeps the user's parentheses (and those of a complement), this repo's Python
    tokenizer normalises them away: the cells must be the same Boolean functions of the senses"""
    g = np.random.default_rng(seed)
    for c in range(len(ofs_a) - 1):
        a, b = tok_a[ofs_a[c]:ofs_a[c + 1]], tok_b[ofs_b[c]:ofs_b[c + 1]]
        if np.array_equal(a, b):
            continue
        for _ in range(512):
            senses = g.integers(0, 2, n_surfaces).astype(bool)
            if _region_truth(list(a), senses) != _region_truth(list(b), senses):
                return False
    return True


SKIP = {"hash_edge", "tok_cell_ofs", "tok_cells", "fine_window", "n_fine_bins", "fine_inv_spacing", "reserved_ce", "mat_base",
        "n_collision_tallies", "n_hex_lattices"}       # derived by omcgpu_init, ignored on input


@pytest.mark.parametrize("name", list(rr.RUNS))
def test_cpp_flattener_equals_python_flattener(dropin, name):
    m, lib, model, box, of, ce = rr.build(name)
    import copy
    m2 = copy.deepcopy(m)
    m2.settings["material_cell_offsets"] = True      # the reference's default (src/settings.cpp)
    py = nuclide.build_ce_model(m2, lib) if ce else mgxs.build_mg_model(m2, lib)
    d = _inputs(name)
    try:
        out = os.path.join(d, "flat.npz")
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "flatten_worker.py"), GPU_LIB, out],
                             cwd=d, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-800:] + res.stderr[-1500:]
        cpp = dict(np.load(out))
    finally:
        shutil.rmtree(d, ignore_errors=True)
    assert "__error__" not in cpp, cpp.get("__error__")
    checked = 0
    for fname, (dt, is_ptr) in abi.model_field_types().items():
        if fname in SKIP:
            continue
        if not ce and (fname.startswith(("nuc_", "rx_", "prod_", "fn_")) or fname in (
                "n_nuclides", "n_log_bins", "log_spacing", "n_reactions", "n_products", "n_functions", "nuct_ofs",
                "n_nuclide_temps")):
            continue
        if ce and fname.startswith(("mg", "sc_", "blk_")):
            continue
        if not is_ptr:
            a, b = cpp[fname].item(), py[fname]
            assert a == b or (isinstance(b, float) and abs(a - b) <= 1e-15 * abs(b)), (fname, a, b)
        else:
            a, b = cpp[fname], py.arrays[fname]
            if fname in ("region_tokens", "cell_region_ofs"):
                assert _regions_equivalent(cpp["region_tokens"], cpp["cell_region_ofs"], py.arrays["region_tokens"],
                                           py.arrays["cell_region_ofs"], int(cpp["n_surfaces"]))
                checked += 1
                continue
            # the Python side pads empty arrays with one zero
            if a.size == 0:
                assert b.size <= 1, fname
                continue
            assert a.size == b.size, (fname, a.size, b.size)
            if dt == np.float64:
                np.testing.assert_allclose(a, b, rtol=1e-15, atol=0, err_msg=fname)
            else:
                assert np.array_equal(a, b), fname
        checked += 1
    assert checked > 60
