"""In-tree build of libbtcsflow.so with nvcc for sm_100a (no torch extension machinery, no JIT cache).

csrc/*.cu are compiled to objects in parallel (the tile-kernel instantiations are split over several translation units
for that reason) and linked into one shared library."""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libbtcsflow.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((PKG_DIR.parent / "include").glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> Path:
    """Compile csrc/*.cu into btcs_pnes_optical_flow_b200/libbtcsflow.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    flags = [*NVCC_FLAGS, *(extra_flags or [])]
    if verbose:
        flags += ["-Xptxas", "-v"]
    with tempfile.TemporaryDirectory(prefix="btcsflow_build_") as tmpdir:
        def compile_one(src: Path):
            obj = Path(tmpdir) / (src.stem + ".o")
            cmd = [nvcc, *flags, "-c", "-o", str(obj), str(src)]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed ({res.returncode}):\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
            return obj, res.stderr

        with ThreadPoolExecutor(max(1, min(len(sources()), os.cpu_count() or 1))) as pool:
            results = list(pool.map(compile_one, sources()))
        if verbose:
            for _, log in results:
                print(log)
        tmp = LIB_PATH.with_suffix(".so.tmp")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(tmp), *(str(o) for o, _ in results)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed ({res.returncode}):\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
