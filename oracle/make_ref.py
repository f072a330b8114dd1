"""Stage the reference's OWN hot-path module files under oracle/_ref/ (git-ignored, NOT gpurun-ignored) so that the
GPU box -- which has no /root/reference -- can time and check against the reference's unmodified code on its host cores.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  The reference is pure Python: there
is nothing to compile, "building" it means placing the files where oracle/ref_import.py can import them behind
oracle/pyg_stub.py (the pure-torch stand-in for the absent torch_geometric / torch_scatter / torch_cluster wheels).
Nothing is copied into the tracked tree: oracle/_ref/ is listed in .gitignore and is produced by this recipe from the
sources where they lie.

    python -m oracle.make_ref            # /root/reference -> oracle/_ref/   (run by __graft_entry__.build() when present)

Files staged (the modules on the path, SURVEY.md section 8a, plus what they import): models/gcn_lib/**, models/multilevel_gnn.py,
models/deepergcn.py, models/diff_pooling.py, models/vae.py, models/utils.py, utils/data_util.py, utils/pyg_util.py, opt.py, config/*.yaml.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MLG_REFERENCE_SRC", "/root/reference")

FILES = ["models/multilevel_gnn.py", "models/deepergcn.py", "models/diff_pooling.py", "models/vae.py", "models/utils.py",
         "utils/data_util.py", "utils/pyg_util.py", "opt.py"]
TREES = ["models/gcn_lib", "config"]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "models", "gcn_lib")):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    for rel in TREES:
        shutil.copytree(os.path.join(SRC, rel), os.path.join(DST, rel),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(SRC + "\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print("staged %d reference files under %s" % (n, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
