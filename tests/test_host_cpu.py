"""CPU tests of the C++ host side (gp_ss_ak_b200/host): the L-BFGS driver replayed against the reference's recorded
probe trace, the data reader / symmetric standardisation, the Statistics and train_model files.  No GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gp_ss_ak_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host_built():
    subprocess.check_call(["make", "-s", "-C", HOST])
    return HOST


def _parse(text):
    rec = {}
    for line in text.splitlines():
        tok = line.split()
        if len(tok) >= 3 and tok[1].isdigit() and tok[2].isdigit():
            r, c = int(tok[1]), int(tok[2])
            rec[tok[0]] = np.array(tok[3:], dtype=float).reshape((c, r)).T
    return rec


@pytest.mark.parametrize("fixture", ["ref_n300.npz", "ref_n2000.npz", "ref_rock_n1000.npz", "ref_exp_n1000.npz", "ref_rbf_n1000.npz"])
def test_lbfgs_replays_reference_trace(host_built, tmp_path, fixture):
    """Every ObjVal / Grad_Values probe of a 30-iteration fit by the UNMODIFIED reference (recorded in
    tests/golden/ref_n300.npz; ref_n2000.npz = BASELINE.json configs[0], 4 iterations / 63 probes) must be requested by the
    host driver in the same order, of the same kind, at the same theta (1e-12 relative); it is answered with the recorded
    f and g.  Exercises cauchy_point, Primal_Conjugate_grad, Efficient_line_search and the memory updates, quirks included."""
    z = np.load(os.path.join(GOLD, fixture))
    trace = tmp_path / "trace.txt"
    with open(trace, "w") as f:
        for k in range(len(z["probe_f"])):
            g = np.nan_to_num(z["probe_g"][k], nan=0.0)
            f.write("%d " % int(z["probe_kind"][k]) + " ".join("%.17g" % v for v in z["probe_theta"][k]) + " %.17g " % z["probe_f"][k]
                    + " ".join("%.17g" % v for v in g) + "\n")
    npar = z["probe_theta"].shape[1]                       # 10: Hyb{ExpAns, Bias}; 4 / 5: Hyb{Exp | RBF, Bias}
    out = subprocess.run([os.path.join(host_built, "tests", "replay_lbfgs"), str(trace), str(int(z["lbfgs_iters"])), "1e-12", "LBFGS", str(npar)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "REPLAY OK probes %d of %d" % (len(z["probe_f"]), len(z["probe_f"])) in out.stdout
    final = np.array(out.stdout.split("final")[1].split(), dtype=float)
    assert np.abs(final - z["theta_fit"].reshape(-1)).max() <= 1e-12


def test_lbfgs_replay_detects_a_wrong_decision(host_built, tmp_path):
    """The harness is not vacuous: perturbing one recorded objective value makes the replay diverge."""
    z = np.load(os.path.join(GOLD, "ref_n300.npz"))
    f_mod = z["probe_f"].copy()
    f_mod[1] = -100.0          # the first line-search trial (recorded 259.3) suddenly beats f0 = -52.5
    trace = tmp_path / "trace.txt"
    with open(trace, "w") as f:
        for k in range(len(f_mod)):
            g = np.nan_to_num(z["probe_g"][k], nan=0.0)
            f.write("%d " % int(z["probe_kind"][k]) + " ".join("%.17g" % v for v in z["probe_theta"][k]) + " %.17g " % f_mod[k]
                    + " ".join("%.17g" % v for v in g) + "\n")
    out = subprocess.run([os.path.join(host_built, "tests", "replay_lbfgs"), str(trace), "30", "1e-12"], capture_output=True, text=True)
    assert out.returncode != 0 and "MISMATCH" in out.stdout


@pytest.mark.parametrize("fixture", ["ref_n300.npz", "ref_n2000.npz"])
def test_reader_standardisation_and_files_match_reference(host_built, tmp_path, fixture):
    z = np.load(os.path.join(GOLD, fixture))
    (tmp_path / "train.txt").write_text(str(z["train_file_text"]))
    (tmp_path / "test.txt").write_text(str(z["test_file_text"]))
    th = z["theta_fit"].reshape(-1)
    (tmp_path / "theta.txt").write_text(" ".join("%.17g" % v for v in th))
    out = subprocess.run([os.path.join(host_built, "tests", "host_io_check"), str(tmp_path / "train.txt"), str(tmp_path / "test.txt"),
                          str(tmp_path / "theta.txt"), str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = _parse(out.stdout)
    assert np.array_equal(rec["X_raw"], z["X_raw"]) and np.array_equal(rec["y_raw"], z["y_raw"])      # atof of %.17g text
    assert np.array_equal(rec["Xs"], z["Xs"]) and np.array_equal(rec["ys"], z["ys"])                  # Control.cpp:299-324
    assert np.array_equal(rec["params"], z["params"])
    assert np.array_equal(rec["Xt"], z["Xt"])                                                         # test mode reuses saved statistics
    assert np.abs(rec["Xt_back"] - z["Xt_raw"]).max() < 1e-9                                          # postData inverts it
    # the reference's hard-coded starting point (Kernel.cpp:763-773, 317-320; GP_Utils.cpp:43)
    assert np.allclose(rec["theta_default"].reshape(-1), [np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016], rtol=0, atol=0)
    # files: byte-identical to what the reference wrote for the same numbers
    assert (tmp_path / "host_model_Statistics.txt").read_text() == str(z["statistics_file_text"])
    assert (tmp_path / "host_model").read_text() == str(z["model_file_text"])
    # read-back: 6 significant digits survive (default ostream precision), structure intact
    rb = rec["theta_readback"].reshape(-1)
    assert np.abs(rb - th).max() <= 5e-6 * np.abs(th).max()
    assert "readback numData %d inputDim 3 outputDim 1 kernel Hyb nkern_params 9" % z["X_raw"].shape[0] in out.stdout


def test_reader_quirks(host_built, tmp_path):
    """Comma or tab separated, '#' comment lines skipped, an empty line counts as a (zero) row (Control.cpp:100-103)."""
    (tmp_path / "a.txt").write_text("# header\n1,2,3,10\n4\t5\t6\t20\n\n7,8,9,30\n# trailing comment\n")
    (tmp_path / "theta.txt").write_text(" ".join(["1"] * 10))
    out = subprocess.run([os.path.join(host_built, "tests", "host_io_check"), str(tmp_path / "a.txt"), str(tmp_path / "a.txt"),
                          str(tmp_path / "theta.txt"), str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = _parse(out.stdout)
    assert rec["X_raw"].shape == (4, 3)
    assert np.array_equal(rec["X_raw"], [[1, 2, 3], [4, 5, 6], [0, 0, 0], [7, 8, 9]])
    assert np.array_equal(rec["y_raw"].reshape(-1), [10, 20, 0, 30])
    # the three spatial columns share one centre / half-range from the global extremes (0 and 9)
    assert np.array_equal(rec["params"][1:], [[4.5, 4.5]] * 3)
    # integers are written as integers in the model file (Kernel.cpp:31-35)
    assert "1 1 1 1 1 1 1 1 \n" in (tmp_path / "host_model").read_text()


def test_streaming_reader_number_parsing_equals_atof(host_built, tmp_path):
    """The streaming reader's in-place number parser (Clinger fast path + atof fallback, host/Control.cpp) must return exactly
    what atof returns for every field: short decimals, exponents, signs, 17-digit values, leading blanks, trailing text, text
    without digits, a last line without newline, CRLF line ends, empty fields (skipped, Control.cpp:52-77)."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.atof.restype = ctypes.c_double
    rng = np.random.default_rng(3)
    fields = ["1", "-2.5", "+.75", "3.", "1e-3", "2.5E+3", "-7e22", "1e23", "123456789012345678", "0.1", "9007199254740993", "1e-22", "1e-23",
              "  7", "12abc", "abc", "-", ".", "1e", "1e+", "0x10", "inf", "nan", "4.9e-324", "1.7976931348623157e308", "00012.500", "-0.0"]
    fields += ["%.17g" % v for v in rng.standard_normal(40) * 10.0 ** rng.integers(-8, 9, 40)]
    fields += ["%.6g" % v for v in rng.standard_normal(40) * 10.0 ** rng.integers(-3, 4, 40)]
    fields += ["%d" % v for v in rng.integers(-10 ** 9, 10 ** 9, 20)]
    rows = [fields[i:i + 4] for i in range(0, len(fields) - 3, 4)]
    text = "# numbers\n" + "\n".join("\t".join(r[:2]) + ",," + ",".join(r[2:]) for r in rows[:-1]) + "\r\n" + "\t".join(rows[-1])   # no final newline
    (tmp_path / "n.txt").write_text(text)
    (tmp_path / "theta.txt").write_text(" ".join(["1"] * 10))
    out = subprocess.run([os.path.join(host_built, "tests", "host_io_check"), str(tmp_path / "n.txt"), str(tmp_path / "n.txt"),
                          str(tmp_path / "theta.txt"), str(tmp_path)], capture_output=True, text=True)
    rec = _parse(out.stdout.split("Xs ")[0])                   # X_raw / y_raw are dumped before the standardisation (inf / nan there)
    got = np.concatenate([rec["X_raw"], rec["y_raw"].reshape(-1, 1)], axis=1)
    want = np.array([[libc.atof(f.encode()) for f in r] for r in rows])
    assert got.shape == want.shape == (len(rows), 4)
    assert np.array_equal(got, want, equal_nan=True), (got - want)


@pytest.mark.parametrize("opt", ["BFGS", "SCG"])
def test_bfgs_and_scg_replay_reference_traces(host_built, tmp_path, opt):
    """The reference's other two optimisers (BFGS is its CLI default, gp_ss_ak.cpp:91), restated in host/Opt_pars.cpp, must
    request exactly the probes the UNMODIFIED reference requested (tests/golden/ref_opt_traces.npz, made by
    make_ref_opt_traces.py): same order, same kind, same theta, and end at the same parameters."""
    z = np.load(os.path.join(GOLD, "ref_opt_traces.npz"))
    kind, theta, fv, gv = z[opt + "_kind"], z[opt + "_theta"], z[opt + "_f"], z[opt + "_g"]
    assert len(fv) >= 12
    trace = tmp_path / "trace.txt"
    with open(trace, "w") as f:
        for k in range(len(fv)):
            g = np.nan_to_num(gv[k], nan=0.0)
            f.write("%d " % int(kind[k]) + " ".join("%.17g" % v for v in theta[k]) + " %.17g " % fv[k] + " ".join("%.17g" % v for v in g) + "\n")
    out = subprocess.run([os.path.join(host_built, "tests", "replay_lbfgs"), str(trace), str(int(z[opt + "_iters"])), "1e-10", opt],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "REPLAY OK probes %d of %d" % (len(fv), len(fv)) in out.stdout
    final = np.array(out.stdout.split("final")[1].split(), dtype=float)
    assert np.abs(final - z[opt + "_theta_fit"].reshape(-1)).max() <= 1e-10


@pytest.mark.parametrize("rows", [5, 40000])
def test_threaded_predict_table_writer_writes_the_reference_loops_bytes(host_built, tmp_path, rows):
    """`<model>_predict.txt` (gp_ss_ak.cpp:470-481: every value inserted into an ofstream, a tab after each, a line per row) is written by
    several host threads since round 2 (Control::writePredictTable; 10 M rows are BASELINE configs[3]).  The bytes must equal the
    reference loop's for every thread count -- magnitudes 1e-9..1e9, both signs, exact integers, +-0, inf and nan included.  The same
    harness checks Control::sortedOrder (rows by ascending observed value, gp_ss_ak.cpp:434-436) against a stable sort_index on a column full of ties."""
    out = subprocess.run([os.path.join(host_built, "tests", "predict_writer_check"), str(rows), str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and "identical 1" in out.stdout and "sorted order identical 1" in out.stdout, out.stdout + out.stderr
    lines = (tmp_path / "w7.txt").read_text().splitlines()
    assert len(lines) == rows + 1 and lines[0] == "# SampleNo, Y,  Yh, StdYh, Inputs" and all(l.endswith("\t") for l in lines[1:4])
