"""In-tree build of the CUDA library (sm_100a only) and the corpus helper."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdlz4_b200.so")
CORPUS_LIB = os.path.join(HERE, "tools", "libdlz4_corpus.so")


def build(force=False):
    if force:
        subprocess.check_call(["make", "-C", CSRC, "-s", "clean"])
    subprocess.check_call(["make", "-C", CSRC, "-s"])
    return LIB
