"""Stage the UNMODIFIED reference sources of the swarm hot path under oracle/_ref/ (container only).

    python -m oracle.make_ref

TEST / BENCH INFRASTRUCTURE.  The reference is pure Python, so "building" it is copying the two
NumPy-only files the path consists of, byte for byte, from where they lie under /root/reference:

    fed_gym/envs/multiagent.py              SwarmEnv                (reference :7-115)
    fed_gym/agents/state_processors.py      SwarmStateProcessor     (reference :15-42)

into oracle/_ref/ with the same relative paths.  oracle/_ref/ is git-ignored (reference sources never
enter this repository's history) but NOT gpurun-ignored: like a built .so it travels to the GPU box,
where /root/reference does not exist, so that bench.py's CPU arm times the reference itself
(``cpu_baseline.kind == "reference"``) and oracle/ref_loader.py can load it there.  A SHA-256 manifest
records what was staged.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
FILES = ["fed_gym/envs/multiagent.py", "fed_gym/agents/state_processors.py"]


def stage(reference_root="/root/reference", quiet=False):
    if not os.path.isfile(os.path.join(reference_root, FILES[0])):
        if not quiet:
            print("reference not mounted at %s: nothing staged" % reference_root)
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(REF, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference_root, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(REF, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files under %s" % (len(FILES), REF))
    return True


def staged():
    return all(os.path.isfile(os.path.join(REF, rel)) for rel in FILES)


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
