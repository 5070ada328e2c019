#!/usr/bin/env python3
"""Regenerate tests/golden/stats_reference.npz -- outputs of the REFERENCE's OWN `stats_by_channel`
(/root/reference/python_src/plot_scripts/multi_campaign_plots_and_ana.py:155-178) on seeded inputs with
NaN gaps.  The plotting script cannot be imported here (PIL, xarray, matplotlib, module-level I/O), so
the function's source is cut out of the file with `ast` and executed as it stands.  `/root/reference`
is read only HERE; tests read the .npz.
"""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/python_src/plot_scripts/multi_campaign_plots_and_ana.py"


def main():
    text = open(SRC).read()
    fn = next(n for n in ast.parse(text).body if isinstance(n, ast.FunctionDef) and n.name == "stats_by_channel")
    ns = {"np": np, "n_chans": 14}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    rng = np.random.default_rng(20261020)
    cases = {}
    for c, (T, nc, nanv, nanr) in enumerate([(521, 14, 0.0, 0.0), (521, 14, 0.03, 0.02), (48, 14, 0.2, 0.1),
                                             (7, 14, 0.4, 0.0), (300, 10, 0.05, 0.05)]):
        v = rng.normal(150, 40, (T, nc))
        r = v + rng.normal(0.4, 1.2, (T, nc))
        v[rng.uniform(size=v.shape) < nanv] = np.nan
        r[rng.uniform(size=r.shape) < nanr] = np.nan
        with np.errstate(all="ignore"):
            sd, bias, rmse = ns["stats_by_channel"](v, r, nc)
        cases.update({"v%d" % c: v, "r%d" % c: r, "sd%d" % c: np.asarray(sd), "bias%d" % c: np.asarray(bias),
                      "rmse%d" % c: np.asarray(rmse)})
    cases["n"] = np.array(5)
    np.savez_compressed(os.path.join(HERE, "stats_reference.npz"), **cases)
    print("ok")


if __name__ == "__main__":
    main()
