"""Classic (monolithic) convex relaxation of the GCS shortest-path problem — Drake-free restatement of the
reference script of the same name.

    python classic_solver.py --test_file <module name in test_data/> [--show_plot <anything>]

Same landmarks on stdout (banner, ``V:`` / ``E:``, ``Beginning MICP Solve.``, ``Solve Time:``, ``Solved using:``,
``Optimal Cost Pre-rounding (Path Length):``, the POST-ROUNDING block) and the same pickle
(``benchmark_data/classic_solver_<test>.pkl``, ``ADMM=False``) as reference ``classic_solver.py:27-228``.  The
program is solved on the host by ``gcs_admm_b200.classic`` (sparse interior-point method); it is the CPU comparator
of the ADMM path, not part of it.
"""
import argparse
import os
import sys

import numpy as np

np.set_printoptions(edgeitems=30, linewidth=250, precision=4, suppress=True)

DEFAULT_TEST_FILE = "benchmark2"


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--test_file", type=str, default=DEFAULT_TEST_FILE,
                        help="The name of the test file (in `test_data` folder) to use (e.g., 'benchmark2').")
    parser.add_argument("--show_plot", type=str, default=True, help="Whether to display plot.")
    parser.add_argument("--seed", type=int, default=None, help="seed of the rounding walk (reference: unseeded)")
    args = parser.parse_args(argv)

    print("=======================================================================")
    print(f"Running Classic Solver on {args.test_file}")
    print("=======================================================================\n")

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import utils
    from gcs_admm_b200.problem_io import load_test_file
    from gcs_admm_b200.classic import solve_classic
    test_data_path = os.path.join(here, "test_data")
    try:
        As, bs, n = load_test_file(args.test_file, test_data_path)
    except ModuleNotFoundError:
        print(f"Error: Test file '{args.test_file}' not found in {test_data_path}.")
        sys.exit(1)

    V, E, I_v_in, I_v_out = utils.build_graph(As, bs)
    print(f"V: {V}")
    print(f"E: {E}")
    print("Beginning MICP Solve.")
    res = solve_classic(As, bs, n, graph=(V, E, I_v_in, I_v_out), seed=args.seed)
    print(f"Solve Time: {res['solve_time']}")
    print("Solved using: gcs_admm_b200.conic (sparse primal-dual interior point)")
    if res["status"] != "optimal" and max(res["residuals"]["pres"], res["residuals"]["dres"]) > 1e-6:
        print("solve failed.")
        print(res["status"], res["residuals"])
        return res
    x_v_sol, y_v_sol, y_e_sol = res["x_v_sol"], res["y_v_sol"], res["y_e_sol"]
    print(f"Optimal Cost Pre-rounding (Path Length): {res['cost']}\n")
    print(f"{x_v_sol=}\n")
    print(f"{y_v_sol=}\n")
    print(f"{y_e_sol=}\n")
    x_v_rounded, y_v_rounded = res["x_v_rounded"], res["y_v_rounded"]
    print("===============================================================")
    print("POST-ROUNDING")
    print("===============================================================")
    print(f"{x_v_rounded=}\n")
    print(f"{y_v_rounded=}\n")
    if args.show_plot == True:  # noqa: E712  (reference semantics)
        utils.visualize_results(As, bs, x_v_sol, y_v_sol, x_v_rounded, y_v_rounded)
    utils.save_data(os.path.join(here, f"benchmark_data/classic_solver_{args.test_file}.pkl"), As, bs, res["solve_time"],
                    res["cost"], x_v_sol, y_v_sol, x_v_rounded, y_v_rounded, ADMM=False)
    return res


if __name__ == "__main__":
    main()
