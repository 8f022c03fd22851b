"""Shared model builders for the tests: the reference's self-contained MG
regression cases with their golden results (results_true.dat)."""
from openmc_b200 import examples, mgxs

# name -> (builder, golden k-combined mean, std, golden tallies [(sum, sum_sq), ...])
GOLDEN = {
    # tests/regression_tests/mg_legendre/results_true.dat
    "mg_legendre": ("1.009220E+00", "9.571832E-03", []),
    # tests/regression_tests/mg_max_order/results_true.dat
    "mg_max_order": ("1.009220E+00", "9.571832E-03", []),
    # tests/regression_tests/mg_survival_biasing/results_true.dat
    "mg_survival_biasing": ("9.889968E-01", "9.144186E-03", []),
    # tests/regression_tests/mg_temperature_multi/results_true.dat
    "mg_temperature_multi": ("1.309371E+00", "6.765039E-03",
                             [("2.532303E+01", "1.282689E+02"), ("9.336894E+01", "1.743765E+03")]),
}


def build(name):
    """-> (ModelHandle, source box, only_fissionable); 1000 particles, 10 batches, 5 inactive."""
    if name == "mg_legendre":
        m, box, of = examples.slab_mg()
        return mgxs.build_mg_model(m, examples.slab_2g_library(1)), box, of
    if name == "mg_max_order":
        m, box, of = examples.slab_mg()
        m.settings["max_order"] = 1
        return mgxs.build_mg_model(m, examples.slab_2g_library(2)), box, of
    if name == "mg_survival_biasing":
        m, box, of = examples.slab_mg()
        m.settings["survival_biasing"] = 1
        return mgxs.build_mg_model(m, examples.slab_2g_library(1)), box, of
    if name == "mg_temperature_multi":
        m, lib, box, of = examples.mg_temperature_multi()
        return mgxs.build_mg_model(m, lib), box, of
    raise KeyError(name)


def fmt(x):
    return "%.6E" % x
