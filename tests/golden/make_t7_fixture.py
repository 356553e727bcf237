"""Copies the reference's Torch7 data fixture used to pin bot7_b200/t7.py.

examples/data/iris_test30.t7 (reference repository, 6528 bytes: a table of four DoubleTensors written
by torch.save) is the only binary-format artefact the reference ships; it is data, not source.  The
reader must parse it and the writer must reproduce it byte for byte (tests/test_t7.py).
Run (where /root/reference is mounted): python tests/golden/make_t7_fixture.py
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/examples/data/iris_test30.t7"

if __name__ == "__main__":
    dst = os.path.join(HERE, "ref_iris_test30.t7")
    shutil.copyfile(SRC, dst)
    os.chmod(dst, 0o644)
    print(dst, hashlib.sha256(open(dst, "rb").read()).hexdigest())
