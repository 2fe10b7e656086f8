#!/usr/bin/env python
"""Recipe for oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE -- never imported by the product package).

The reference is pure Python with no packaging, and /root/reference does not exist on the GPU box.  This script packs the
UNMODIFIED reference modules the hot path needs (models/*.py and the dist.py they import) into ONE archive,

    oracle/_ref/reference_path.tar.gz        (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so)

so that `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's OWN functions
(VAR.autoregressive_infer_cfg models/var.py:128, SDVAR.sdvar_autoregressive_infer_cfg_sd_test3 models/var.py:605) on the
box's host cores.  No reference source enters the repository's history; oracle/ref_runtime.py unpacks the archive into a
temporary directory at run time.  `__graft_entry__.build()` runs this whenever /root/reference is present."""
import io
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SDVAR_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "reference_path.tar.gz")
MEMBERS = ["dist.py"] + [os.path.join("models", f) for f in ("__init__.py", "basic_vae.py", "basic_var.py", "helpers.py", "quant.py",
                                                               "var.py", "vqvae.py")]


def build(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(REF, "models")):
        return OUT if os.path.exists(OUT) else ""
    srcs = [os.path.join(REF, m) for m in MEMBERS]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for m, s in zip(MEMBERS, srcs):
            tar.add(s, arcname=m)
    with open(OUT, "wb") as f:
        f.write(buf.getvalue())
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv) or "reference not present and no archive built earlier")
