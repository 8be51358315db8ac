"""Install the UNMODIFIED reference under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

    python baseline/install_reference.py            # in the build container (needs /root/reference)

1. Tries the contract's `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
   /root/reference`.  The reference is a source tree of scripts with no setup.py / pyproject.toml, so pip refuses it
   ("does not appear to be a Python project"); the outcome is written to baseline/_ref/INSTALL_LOG.txt.
2. Falls back to what the reference's own README does — run from a checkout — by copying the importable packages the hot
   path needs (`src/`, `core/`, `backend_config.py`) byte for byte.  Nothing is edited; bench.py applies the three offline
   monkey-patches of SURVEY.md App. B at run time (HF from_pretrained -> random-init config, stub tokenizer, torchvision
   vit_b_16(weights=None)), because the GPU box has no network, no timm and no tokenizer files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("VC_REFERENCE", "/root/reference"))
DST = ROOT / "baseline" / "_ref"
PARTS = ["src", "core", "backend_config.py"]


def main() -> int:
    if not REF.exists():
        print(f"{REF} not found: nothing to install (the GPU box uses the prebuilt baseline/_ref)")
        return 0 if DST.exists() else 1
    DST.mkdir(parents=True, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
           "--target", str(DST), str(REF)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = [f"$ {' '.join(cmd)}", f"rc={r.returncode}", r.stdout[-2000:], r.stderr[-2000:]]
    if r.returncode != 0:
        log.append("pip cannot install the reference (no setup.py / pyproject.toml): copying the importable source tree instead")
        for part in PARTS:
            s, d = REF / part, DST / part
            if s.is_dir():
                shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            elif s.exists():
                shutil.copy2(s, d)
    (DST / "INSTALL_LOG.txt").write_text("\n".join(log))
    n = sum(1 for _ in DST.rglob("*.py"))
    print(f"baseline/_ref ready: {n} python files ({'pip' if r.returncode == 0 else 'source copy'})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
