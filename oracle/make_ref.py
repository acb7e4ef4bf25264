"""Recipe that stages the REFERENCE'S OWN modules of the hot path under ``oracle/_ref/`` -- TEST / BASELINE INFRASTRUCTURE.

    python oracle/make_ref.py            (run by __graft_entry__.build() whenever /root/reference is present)

``/root/reference`` does not exist on the GPU box; ``oracle/_ref/`` is git-ignored (no reference source ever enters the
history) but travels with the repository snapshot, like the built ``.so``.  What is staged, unmodified:

    hf_hypernet/configuration_hypernet.py, hf_hypernet/modeling_hypernet.py   the PyTorch binding of the hypernetwork
    zett/utils.py (+ data/madlad400_metadata.csv, read at import)             get_surface_form_matrix + CHARS_TO_BYTES

``oracle/ref_loader.py`` imports them with the two offline stubs SURVEY.md appendix A describes.  Consumers: ``bench.py
--impl reference`` / the ``cpu_baseline`` leg (``kind: "reference"``) and tests that compare the restatements in
``oracle/`` with the real thing.  The product (``zett_b200/``) never imports any of it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ZETT_REFERENCE_DIR", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = ["hf_hypernet/__init__.py", "hf_hypernet/configuration_hypernet.py", "hf_hypernet/modeling_hypernet.py", "zett/utils.py",
         "data/madlad400_metadata.csv"]   # zett/utils.py reads the csv at import (zett/utils.py:28)


def main() -> int:
    if not os.path.isdir(REFERENCE):
        print("make_ref: %s not present (GPU box?) -- keeping whatever oracle/_ref already holds" % REFERENCE)
        return 0
    staged = []
    for rel in FILES:
        src = os.path.join(REFERENCE, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
            staged.append(rel)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
    init = os.path.join(DEST, "zett", "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()   # the reference's zett/__init__.py imports the whole JAX training stack; not needed
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write("staged unmodified from %s by oracle/make_ref.py:\n%s\n" % (REFERENCE, "\n".join(staged)))
    print("make_ref: staged %d files under %s" % (len(staged), DEST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
