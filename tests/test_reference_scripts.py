"""The reference's own callers, unmodified, on this repo's `dpf_cpp` (north_star: "drops into
sample.py and benchmark.py unchanged").

`__graft_entry__.build()` stages /root/reference/{dpf.py,sample.py,benchmark.py} byte for byte
into the git-ignored oracle/_ref/scripts*/ (they travel to the GPU box like the prebuilt
reference .so files).  Each script runs in a subprocess whose only link to this repo is
PYTHONPATH=gpu-dpf_b200:
  * scripts/      holds the reference's dpf.py too, so `import dpf` is the REFERENCE module and only
                  `dpf_cpp` (the pybind boundary, dpf_wrapper.cu:188-204) is ours;
  * scripts_api/  holds sample.py + benchmark.py only, so `import dpf` is our dpf.py.
"""
import hashlib
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-dpf_b200")
SCRIPTS = os.path.join(ROOT, "oracle", "_ref", "scripts")
SCRIPTS_API = os.path.join(ROOT, "oracle", "_ref", "scripts_api")
OUT = os.path.join(ROOT, "gpurun_out")


def _run(cwd, script, timeout=900):
    env = dict(os.environ, PYTHONPATH=PKG, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, script], cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, "%s failed:\n%s\n%s" % (script, r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


def _need(path):
    if not os.path.isfile(os.path.join(path, "sample.py")):
        pytest.skip("reference scripts not staged (oracle/_ref/scripts*: run __graft_entry__.build() where /root/reference exists)")


def _save(name, text):
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)
    except OSError:
        pass


def test_staged_scripts_are_the_reference_files():
    """Staging is a byte-for-byte copy (checked wherever the reference tree is visible)."""
    _need(SCRIPTS)
    if not os.path.isfile("/root/reference/dpf.py"):
        pytest.skip("reference tree not on this machine")
    for name in ("dpf.py", "sample.py", "benchmark.py"):
        a = hashlib.sha256(open(os.path.join("/root/reference", name), "rb").read()).hexdigest()
        b = hashlib.sha256(open(os.path.join(SCRIPTS, name), "rb").read()).hexdigest()
        assert a == b, name


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["scripts", "scripts_api"])
def test_sample_py(where):
    """sample.py:39-56: two-server PIR of entry 42; it asserts rec == 42 itself and prints `a b rec`."""
    cwd = os.path.join(ROOT, "oracle", "_ref", where)
    _need(cwd)
    out = _run(cwd, "sample.py")
    last = out.strip().splitlines()[-1].split()
    assert len(last) == 3 and int(last[2]) == 42 and int(last[0]) - int(last[1]) == 42, out
    _save("r2_sample_py_%s.txt" % where, out)


@pytest.mark.gpu
def test_reference_dpf_py_self_tests():
    """dpf.py:359-367 __main__: test_cpu_dpf, test_cpu_dpf_one_hot, test_gpu_dpf, test_gpu_dpf_nopad,
    test_gpu_dpf_sweep, test_gpu_dpf_perf -- the reference module, our extension underneath."""
    _need(SCRIPTS)
    out = _run(SCRIPTS, "dpf.py")
    for needle in ("Pass CPU check", "Pass CPU (one-hot only) check", "Pass GPU check", "Pass GPU (nopad) check",
                   "Pass GPU (sweep) check", "dpfs/sec"):
        assert needle.lower() in out.lower(), (needle, out[-1500:])
    _save("r2_reference_dpf_py_selftests.txt", out)


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["scripts", "scripts_api"])
def test_benchmark_py(where):
    """benchmark.py:4-7: N in {2^14, 2^16, 2^18, 2^20} x {AES128, SALSA20, CHACHA20}: twelve perf lines."""
    cwd = os.path.join(ROOT, "oracle", "_ref", where)
    _need(cwd)
    out = _run(cwd, "benchmark.py", timeout=1500)
    perf = re.findall(r"DPF\(entries=(\d+), entry_size=16, prf_method=(\w+)\) Key Size: 2096 bytes, Perf: (\d+) dpfs/sec", out)
    assert len(perf) == 12, out
    assert [int(p[0]) for p in perf] == [16384] * 3 + [65536] * 3 + [262144] * 3 + [1048576] * 3
    assert [p[1] for p in perf[:3]] == ["AES128", "SALSA20", "CHACHA20"]
    assert all(int(p[2]) > 0 for p in perf)
    _save("r2_benchmark_py_%s.txt" % where, out)
