"""Can the reference's own model arithmetic (matgl @ 5171392 + dgl, /root/reference/pyproject.toml:25-35) be run here to pin
oracle/chgnet_ref.py?  Test infrastructure: probes every route and prints a log (committed under profiles/).

Routes: import from the environment; import from baseline/_ref (driver-provided reference install);
offline `pip download` / `pip install --no-index` from /opt/wheelhouse.  If one of them ever yields an importable matgl,
tests/golden/make_chgnet_golden.py (same directory convention as make_golden.py) is the next step.
"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODS = ["matgl", "dgl", "ase", "pymatgen", "torch", "numpy"]


def try_import(extra_path=None):
    out = {}
    if extra_path and os.path.isdir(extra_path):
        sys.path.insert(0, extra_path)
    for m in MODS:
        try:
            mod = importlib.import_module(m)
            out[m] = f"OK {getattr(mod, '__version__', '?')}"
        except Exception as ex:  # noqa: BLE001
            out[m] = f"FAIL {type(ex).__name__}: {str(ex)[:80]}"
    return out


def main():
    print("# probe of the reference's third-party model code (matgl / dgl) -- can the oracle be pinned?")
    print("python", sys.version.split()[0], "| cwd", os.getcwd())
    print("\n## 1. plain import")
    for k, v in try_import().items():
        print(f"  {k:10s} {v}")
    ref = os.path.join(ROOT, "baseline", "_ref")
    print(f"\n## 2. baseline/_ref present: {os.path.isdir(ref)}")
    if os.path.isdir(ref):
        for k, v in try_import(ref).items():
            print(f"  {k:10s} {v}")
    wh = "/opt/wheelhouse"
    print(f"\n## 3. wheelhouse {wh}: {os.path.isdir(wh)}")
    if os.path.isdir(wh):
        names = sorted(os.listdir(wh))
        hits = [n for n in names if any(t in n.lower() for t in ("matgl", "dgl", "ase-", "pymatgen", "e3nn", "mace"))]
        print(f"  {len(names)} files; matgl/dgl/ase/pymatgen/e3nn/mace wheels: {hits or 'none'}")
    for pkg in ("matgl", "dgl"):
        cmd = [sys.executable, "-m", "pip", "download", "--no-index", "--find-links", wh, "--no-deps", "-d", "/tmp/_probe_dl", pkg]
        r = subprocess.run(cmd, capture_output=True, text=True)
        tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
        print(f"  pip download --no-index {pkg}: rc={r.returncode}  {tail[0][:140]}")
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", wh, "--target", "/tmp/_probe_ref",
           "/root/reference"]
    if os.path.isdir("/root/reference"):
        r = subprocess.run(cmd, capture_output=True, text=True)
        lines = (r.stdout + r.stderr).strip().splitlines()
        print(f"\n## 4. pip install --no-index /root/reference (the reference arm's install recipe): rc={r.returncode}")
        for ln in lines[-4:]:
            print("  " + ln[:160])
        r2 = subprocess.run(cmd[:-1] + ["--no-deps", "/root/reference"], capture_output=True, text=True)
        print(f"   with --no-deps: rc={r2.returncode}  {(r2.stdout + r2.stderr).strip().splitlines()[-1][:140] if (r2.stdout + r2.stderr).strip() else ''}")
        if r2.returncode == 0:
            sys.path.insert(0, "/tmp/_probe_ref")
            try:
                importlib.import_module("DistMLIP.implementations.matgl")
                print("   import DistMLIP.implementations.matgl: OK")
            except Exception as ex:  # noqa: BLE001
                print(f"   import DistMLIP.implementations.matgl: FAIL {type(ex).__name__}: {str(ex)[:100]}")
    print("\n## verdict")
    ok = try_import().get("matgl", "").startswith("OK")
    print("  matgl importable:", ok, "-> the CHGNet layer internals stay RECALLED (parity unpinned)" if not ok else "")


if __name__ == "__main__":
    main()
