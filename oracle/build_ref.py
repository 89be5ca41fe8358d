"""TEST / BASELINE INFRASTRUCTURE ONLY.

Recipe for oracle/_ref/: a verbatim copy of the reference's own Python modules, taken from where they lie
(/root/reference, read-only) into the git-ignored directory oracle/_ref/ so that they travel to the GPU box
(`/root/reference` does not exist there).  Nothing is compiled: the reference is 100 % Python (CC0, LICENSE:1-8).
`bench.py --impl reference` and the cpu_baseline leg import them through oracle/ref_shim.py (time.clock,
asfarray, and a SuperLU-backed stand-in for the absent cvxopt) and drive the UNMODIFIED
fill_alpha_harmonic.fill_alpha_harmonic through its stock path.  The product never imports anything here.

    python oracle/build_ref.py
"""
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["recovery.py", "fill_alpha_harmonic.py", "tictoc.py", "highlighting.py", "poisson.py", "LICENSE"]


def build_ref():
    """Returns True if oracle/_ref/ is (now) populated."""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            s, d = os.path.join(SRC, f), os.path.join(DST, f)
            if os.path.isfile(s) and (not os.path.isfile(d) or os.path.getmtime(d) < os.path.getmtime(s)):
                shutil.copyfile(s, d)
    return os.path.isfile(os.path.join(DST, "recovery.py"))


if __name__ == "__main__":
    print("oracle/_ref ready" if build_ref() else "no reference tree at %s and no earlier copy" % SRC)
