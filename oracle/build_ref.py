"""TEST / BENCH INFRASTRUCTURE — recipe that makes the UNMODIFIED reference's step path runnable on the GPU box.

    python oracle/build_ref.py            # needs /root/reference (build container); writes oracle/_ref/

``/root/reference`` does not exist on the GPU box, and no reference source is ever copied into this repository.  This recipe
COMPILES the handful of reference modules the continuous-env step path imports (the list below, found by importing the path
once and reading ``sys.modules``) with ``py_compile`` from the sources where they lie, and writes only the resulting bytecode
(sourceless layout ``pkg/module.pyc``, packed into ONE archive ``oracle/_ref/cyberbattle_ref.zip`` that Python imports through
zipimport — the snapshot that ships the repo to the GPU box drops loose ``*.pyc`` files) into ``oracle/_ref/`` — git-ignored,
NOT gpurun-ignored, so it travels to the box like the repo's own built ``.so``.  ``bench.py --impl reference`` (and ``cpu_baseline``) then imports ``cyberbattle`` from
there under the stub modules of ``oracle/shims`` and steps the reference's own
``RandomSwitchEnv(envs_list=[CyberBattleCompressedEnv])`` (``cpu_baseline.kind == "reference"``); without ``oracle/_ref`` they
fall back to the oracle port (``kind == "port"``).  ``__graft_entry__.build()`` runs this recipe when the reference is mounted.
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT, "cyberbattle_ref.zip")
SRC = os.environ.get("CBS_REFERENCE_SRC", "/root/reference")
# every reference module `RandomSwitchEnv.step/reset` over a CyberBattleCompressedEnv imports (sys.modules after a rollout),
# plus the scenario generator the envs are built with (Model(network=G) -> generate_network.py)
MODULES = [
    "cyberbattle/__init__.py",
    "cyberbattle/_env/__init__.py",
    "cyberbattle/_env/cyberbattle_env.py",
    "cyberbattle/_env/cyberbattle_env_compressed.py",
    "cyberbattle/_env/cyberbattle_env_switch.py",
    "cyberbattle/_env/static_defender.py",
    "cyberbattle/gae/__init__.py",
    "cyberbattle/gae/model.py",
    "cyberbattle/simulation/__init__.py",
    "cyberbattle/simulation/attacker_actions.py",
    "cyberbattle/simulation/generate_network.py",
    "cyberbattle/simulation/model.py",
    "cyberbattle/simulation/static_defender_actions.py",
    "cyberbattle/utils/classifier_utils.py",
    "cyberbattle/utils/data_utils.py",
    "cyberbattle/utils/encoding_utils.py",
    "cyberbattle/utils/file_utils.py",
    "cyberbattle/utils/gym_utils.py",
    "cyberbattle/utils/networkx_utils.py",
]


def build(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "cyberbattle")):
        raise RuntimeError(f"reference tree not found at {SRC}")
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in MODULES:
            dst = os.path.join(tmp, rel[:-3] + ".pyc")
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            # dfile: the path tracebacks show (the reference's own file); unchecked: the source is not on the box to be compared with
            py_compile.compile(os.path.join(SRC, rel), cfile=dst, dfile=os.path.join("/root/reference", rel), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            z.write(dst, rel[:-3] + ".pyc")
        # cyberbattle/utils has no __init__.py in the reference (a namespace package); zipimport needs a regular package, so an
        # EMPTY module is compiled in its place
        have = {os.path.dirname(rel) for rel in MODULES if rel.endswith("__init__.py")}
        for pkg in sorted({os.path.dirname(rel) for rel in MODULES} - have):
            empty = os.path.join(tmp, "_empty.py")
            open(empty, "w").close()
            dst = os.path.join(tmp, pkg, "__init__.pyc")
            py_compile.compile(empty, cfile=dst, dfile=os.path.join("/root/reference", pkg, "__init__.py"), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            z.write(dst, os.path.join(pkg, "__init__.pyc"))
    with open(os.path.join(OUT, "BUILT_FROM.txt"), "w") as f:
        f.write(f"bytecode of {len(MODULES)} modules compiled from {SRC} by oracle/build_ref.py with python {sys.version.split()[0]}\n")
    if verbose:
        print(f"oracle/_ref: {len(MODULES)} modules compiled from {SRC} -> {os.path.relpath(ARCHIVE, os.path.dirname(HERE))}")
    return ARCHIVE


if __name__ == "__main__":
    build()
