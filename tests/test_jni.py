"""The Java half of the boundary, as far as an image without a JDK can take it:

  * java/carskit/b200/Native.java declares one native method per JNI function jni/carskit_b200_jni.c defines, with
    matching arity (checked textually);
  * the glue compiles warning-free against jni/stub/jni.h, links libcarskit_b200.so, and is DRIVEN by a fake JNIEnv
    (tests/jni_fake_env.c): without a device it must surface CARS_E_NO_DEVICE as a RuntimeException with every critical
    region closed and no JNI call inside one; on the GPU box it trains and must print exactly what examples/c_client.c
    prints for the same inputs;
  * every reference member the Java sources touch exists in the reference's jars with that signature (class-file reader).
"""
import os
import re
import subprocess

import pytest

from carskit_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "carskit_b200")


def build_fake_jvm(tmp_path):
    exe = str(tmp_path / "jni_fake")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "jni", "stub"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "carskit_b200_jni.c"),
                           os.path.join(ROOT, "tests", "jni_fake_env.c"), "-L", LIBDIR, "-lcarskit_b200",
                           f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_native_java_and_jni_glue_declare_the_same_methods():
    java = open(os.path.join(ROOT, "java", "carskit", "b200", "Native.java")).read()
    glue = open(os.path.join(ROOT, "jni", "carskit_b200_jni.c")).read()
    jm = {}
    for m in re.finditer(r"public static native\s+[\w\[\]]+\s+(\w+)\s*\(([^)]*)\)", java, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        jm[m.group(1)] = len(args)
    gm = {}
    for m in re.finditer(r"JNICALL Java_carskit_b200_Native_(\w+)\(([^)]*)\)", glue, flags=re.S):
        gm[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()]) - 2  # JNIEnv*, jclass
    assert jm == gm and len(jm) >= 16


@pytest.mark.skipif(_has_cuda(), reason="on a GPU box the glue trains (test below)")
def test_jni_glue_compiles_links_and_surfaces_no_device_as_an_exception(cars_lib, tmp_path):
    p = subprocess.run([build_fake_jvm(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 3, p.stdout + p.stderr
    assert "deviceCount = 0" in p.stdout and "RuntimeException" in p.stdout and "no CPU path" in p.stdout


@pytest.mark.gpu
def test_jni_glue_trains_like_the_plain_c_client(cars_lib, tmp_path):
    exe = build_fake_jvm(tmp_path)
    cc = str(tmp_path / "c_client")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_client.c"),
                           "-L", LIBDIR, "-lcarskit_b200", f"-Wl,-rpath,{LIBDIR}", "-o", cc])
    a = subprocess.run([exe], capture_output=True, text=True)
    b = subprocess.run([cc], capture_output=True, text=True)
    assert a.returncode == 0, a.stdout + a.stderr
    assert b.returncode == 0, b.stdout + b.stderr
    la = [ln for ln in a.stdout.splitlines() if ln.startswith("iter ")]
    lb = [ln for ln in b.stdout.splitlines() if ln.startswith("iter ")]
    assert la == lb and len(la) == 3
    pa = re.search(r"P\[0\]\[0\] = (\S+),", a.stdout).group(1)
    pb = re.search(r"P\[0\]\[0\] = (\S+),", b.stdout).group(1)
    assert pa == pb


REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "jar", "CARSKit-v0.4.0.jar")), reason="reference jars not on this box")
def test_every_reference_member_the_java_sources_use_exists():
    from tests.tools.classfile import Jar
    jar = Jar(os.path.join(REF, "jar", "CARSKit-v0.4.0.jar"), os.path.join(REF, "lib", "librec-v1.4-alpha.jar"),
              os.path.join(REF, "lib", "happy.coding.utils-1.2.6.jar"))

    def has_method(cls, name, desc=None):
        c = cls
        while c and jar.has(c):
            cf = jar.load(c)
            if any(n == name and (desc is None or d == desc) for (n, d) in cf.methods):
                return True
            c = cf.super_name
        return False

    def field(cls, name):
        c = cls
        while c and jar.has(c):
            cf = jar.load(c)
            if name in cf.fields:
                return cf.fields[name]
            c = cf.super_name
        return None
    sm, dm, dv, dao = "librec/data/SparseMatrix", "librec/data/DenseMatrix", "librec/data/DenseVector", "carskit/data/processor/DataDAO"
    for cls, name, desc in [(sm, "getRowPointers", "()[I"), (sm, "getColumnIndices", "()[I"), (sm, "getData", "()[D"),
                            (dm, "numRows", "()I"), (dm, "numColumns", "()I"), (dm, "get", "(II)D"), (dm, "set", "(IID)V"),
                            (dm, "init", "(DD)V"), (dv, "getData", "()[D"), (dv, "get", "(I)D"), (dv, "init", "()V"),
                            (dao, "getUserIdFromUI", "(I)I"), (dao, "getItemIdFromUI", "(I)I"), (dao, "numContexts", "()I"),
                            (dao, "numContextDims", "()I"), ("carskit/generic/ContextRecommender", "getConditions", "(I)Ljava/util/List;"),
                            ("carskit/generic/IterativeRecommender", "isConverged", "(I)Z"),
                            ("librec/data/SymmMatrix", "get", "(II)D"), ("librec/data/SymmMatrix", "set", "(IID)V"),
                            ("happy/coding/io/LineConfiger", "getInt", "(Ljava/lang/String;I)I"),
                            ("happy/coding/io/LineConfiger", "getFloat", "(Ljava/lang/String;)F"),
                            ("happy/coding/io/LineConfiger", "getString", "(Ljava/lang/String;Ljava/lang/String;)Ljava/lang/String;")]:
        assert has_method(cls, name, desc), (cls, name, desc)
    it = "carskit/generic/IterativeRecommender"
    # protected (0x4) or public (0x1), never private (0x2): reachable from a subclass in another package
    for cls, name, desc in [(it, "P", None), (it, "Q", None), (it, "userBias", None), (it, "itemBias", None), (it, "lRate", "D"),
                            (it, "loss", "D"), (it, "regU", "F"), (it, "regC", "F"), (it, "numFactors", "I"), (it, "numIters", "I"),
                            ("carskit/alg/cars/adaptation/dependent/CAMF", "condBias", None),
                            ("carskit/alg/cars/adaptation/dependent/CAMF", "icBias", None),
                            ("carskit/alg/cars/adaptation/dependent/CAMF", "ucBias", None),
                            ("carskit/alg/cars/adaptation/dependent/CAMF", "ccMatrix_ICS", "Llibrec/data/SymmMatrix;"),
                            ("carskit/alg/cars/adaptation/dependent/CAMF", "cfMatrix_LCS", "Llibrec/data/DenseMatrix;"),
                            ("carskit/alg/cars/adaptation/dependent/CAMF", "cVector_MCS", "Llibrec/data/DenseVector;"),
                            ("carskit/alg/baseline/cf/SVDPlusPlus", "Y", "Llibrec/data/DenseMatrix;"),
                            ("carskit/generic/ContextRecommender", "EmptyContextConditions", "Ljava/util/ArrayList;"),
                            ("carskit/generic/Recommender", "train", "Llibrec/data/SparseMatrix;"),
                            ("carskit/generic/Recommender", "trainMatrix", None), ("carskit/generic/Recommender", "rateDao", None),
                            ("carskit/generic/Recommender", "algoOptions", None), ("carskit/generic/Recommender", "globalMean", "D"),
                            ("carskit/generic/Recommender", "fold", "I"), ("carskit/generic/Recommender", "numUsers", "I"),
                            ("carskit/generic/ContextRecommender", "numConditions", "I"),
                            ("carskit/alg/cars/adaptation/dependent/dev/CAMF_CUCI", "icBias", "Lcom/google/common/collect/Table;")]:
        f = field(cls, name)
        assert f is not None, (cls, name)
        assert not (f[1] & 0x2), f"{cls}.{name} is private"
        if desc:
            assert f[0] == desc, (cls, name, f[0])
    # FM keeps its model private -- why FM_B200 extends ContextRecommender instead of FM
    fm = jar.load("carskit/alg/cars/adaptation/dependent/FM")
    assert all(fm.fields[n][1] & 0x2 for n in ("w0", "w", "V", "regLw", "regLf"))
