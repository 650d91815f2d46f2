"""Generate tests/golden/api_surface.npz: outputs and gradients of the example-API-surface cases
(tests/test_host_vs_reference_cpu.py: CASES) computed by the UNMODIFIED reference on its CPU path.
Run HERE only (needs /root/reference):  python oracle/make_api_surface_golden.py"""
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    spec = importlib.util.spec_from_file_location("cases_mod", os.path.join(ROOT, "tests", "test_host_vs_reference_cpu.py"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "numpy-nn-model_b200"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = os.path.join(ROOT, "tests", "golden", "api_surface.npz")
    with tempfile.TemporaryDirectory() as d:
        cases = os.path.join(d, "cases.py")
        open(cases, "w").write(mod.CASES)
        runner = os.path.join(d, "runner.py")
        open(runner, "w").write(mod.REF_RUNNER % dict(ref=mod.REF, cases=cases, out=out))
        env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
        subprocess.run([sys.executable, runner], check=True, env=env, cwd=d)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
