"""Generates the committed golden fixtures from the reference's own test data.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources:
  tests/MultiIK.zip -> MultiIK/triBlockDiag_{G,C,a,u,sol}.txt   ("Sequential IK",   tests/BlockGISolverTest.in.cpp:172-188)
                    -> MultiIK/arrowAllData.txt                  ("Simultaneous IK", tests/BlockGISolverTest.in.cpp:273-284,
                                                                   format: tests/IKmatReader.cpp:117-170)
The matrices are stored losslessly (float64, deflate); only zeros are dropped by the compression.
"""
import io
import os
import zipfile

import numpy as np

REF = "/root/reference/tests/MultiIK.zip"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_mat(text):
    return np.loadtxt(io.StringIO(text), ndmin=2)


def read_ik_file(text):
    fields = {}
    name = None
    buf = []
    for line in text.splitlines():
        if line.startswith("===="):
            if name is not None:
                fields[name] = "\n".join(buf)
            name = line.strip("= ").strip()
            buf = []
        else:
            buf.append(line)
    fields[name] = "\n".join(buf)
    out = {}
    for k, v in fields.items():
        v = v.strip()
        out[k] = np.loadtxt(io.StringIO(v), ndmin=2) if v else np.zeros((0, 0))
    return out


def main():
    z = zipfile.ZipFile(REF)
    rd = lambda n: z.read("MultiIK/" + n).decode()
    G = read_mat(rd("triBlockDiag_G.txt"))
    Cm = read_mat(rd("triBlockDiag_C.txt"))  # 1621 x 387: one constraint per ROW (the test transposes it)
    a = read_mat(rd("triBlockDiag_a.txt")).ravel()
    u = read_mat(rd("triBlockDiag_u.txt")).ravel()
    sol = read_mat(rd("triBlockDiag_sol.txt")).ravel()
    np.savez_compressed(os.path.join(HERE, "multiik_sequential.npz"), G=G, C=Cm, a=a, u=u, sol=sol)
    d = read_ik_file(rd("arrowAllData.txt"))
    n = int(d["dim_var"][0, 0])
    assert int(d["dim_eq"][0, 0]) == 0 and int(d["dim_ineq"][0, 0]) == 25
    np.savez_compressed(os.path.join(HERE, "multiik_simultaneous.npz"), G=d["Q"], a=d["c"].ravel(), C=d["C"], u=d["d"].ravel(),
                        xl=d["x_min"].ravel(), xu=d["x_max"].ravel())
    print("sequential", G.shape, Cm.shape, a.shape, u.shape, sol.shape)
    print("simultaneous", n, d["Q"].shape, d["C"].shape, d["d"].shape, d["x_min"].shape)
    for f in os.listdir(HERE):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
