"""Pin the machine code of the built library: `python scripts/pin_sass.py` writes tests/golden/sass_r02.txt (hash, two spaces, demangled
kernel name), the fixture of tests/test_abi_cpu.py::test_machine_code_of_the_measured_kernels_is_unchanged.  Run it on the commit whose
numbers go into profiles/ (i.e. after the last GPU measurement of a kernel change)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_hashes.py"), os.path.join(ROOT, "geophyinv.jl_b200", "libgpifdtd.so")],
                     capture_output=True, text=True, check=True).stdout
with open(os.path.join(ROOT, "tests", "golden", "sass_r02.txt"), "w") as f:
    for line in out.splitlines():
        h, _, name = line.split(" ", 2)
        f.write(f"{h}  {name.strip()}\n")
print(f"pinned {len(out.splitlines())} kernels")
