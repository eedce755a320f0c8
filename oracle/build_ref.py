"""Stage the UNMODIFIED reference hot-path sources into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST/BENCH INFRASTRUCTURE.  The reference is pure Python: nothing is compiled, the files are copied byte for
byte from /root/reference so that `bench.py --impl reference` and the `cpu_baseline` leg can EXECUTE the
reference's own functions on the GPU box (which has no /root/reference) — `cpu_baseline.kind = "reference"`.
Nothing here enters the repository history (oracle/_ref/ is in .gitignore) and nothing under litehandnet_b200/
reads it.  Run by __graft_entry__.build() whenever /root/reference is present:

    python -m oracle.build_ref            # copies, then loads the copy through oracle/ref_loader.py as a check
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LHN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")

# the modules oracle/ref_loader.py loads (SURVEY.md §8c load order) + the loss package
FILES = [
    "config/__init__.py",
    "datasets/data_pipeline/post_transforms.py",
    "datasets/data_pipeline/generateTarget.py",
    "datasets/data_pipeline/generate_simder.py",
    "utils/bbox_metric.py",
    "utils/heatmap_post_processing.py",
    "utils/evaluation.py",
    "utils/post_processing/evaluation/top_down_eval.py",
    "utils/post_processing/decoder.py",
    "utils/transforms.py",
    "utils/result_parser.py",
    "utils/SPheatmapParser.py",
    "utils/HeatmapParser.py",
]
DIRS = ["loss"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "utils")):
        return None
    files = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".py"):
                files.append(f"{d}/{name}")
    manifest = []
    for rel in files:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest.append(f"{_sha(dst)}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"staged {len(files)} reference files into {DST}")
    return DST


if __name__ == "__main__":
    dst = build(verbose=True)
    if dst is None:
        sys.exit("reference tree not found")
    os.environ["LHN_REFERENCE_ROOT"] = dst
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    ref = ref_loader.load()
    print("loaded from the staged copy:", sorted(k for k in vars(ref) if not k.startswith("_")))
