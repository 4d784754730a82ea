"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; compute entry
points refuse to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+\**([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", re.M)


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef\s+struct[^;]*?\{.*?\}[^;]*;", "", text, flags=re.S)
    return sorted(set(DECL.findall(text)))


@pytest.mark.parametrize("header", ["ieache_b200.h", "tfhe/tfhe.h", "tfhe/tfhe_io.h"])
def test_every_declared_symbol_is_exported(pkg, header):
    names = declared(header)
    assert len(names) >= 4
    lib = pkg.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_cloud_c_import_table_is_covered(pkg):
    # the symbols the reference's Cloud/cloud binary imports from libtfhe
    needed = ["bootsAND", "bootsXOR", "bootsNOT", "bootsCOPY", "bootsCONSTANT", "bootsSymEncrypt", "bootsSymDecrypt",
              "new_LweSample_array", "delete_LweSample_array", "new_gate_bootstrapping_ciphertext_array",
              "delete_gate_bootstrapping_ciphertext_array", "new_tfheGateBootstrappingCloudKeySet_fromFile",
              "new_tfheGateBootstrappingSecretKeySet_fromFile", "delete_gate_bootstrapping_cloud_keyset",
              "delete_gate_bootstrapping_secret_keyset", "import_gate_bootstrapping_ciphertext_fromFile",
              "export_gate_bootstrapping_ciphertext_toFile", "bootsOR", "bootsMUX", "bootsNAND"]
    lib = pkg.lib()
    assert all(hasattr(lib, n) for n in needed)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.EngineError):
        pkg.Engine(0)


def test_product_never_links_the_oracle():
    import subprocess
    so = os.path.join(ROOT, "ie-ache_b200", "libieache_b200.so")
    out = subprocess.run(["nm", "-D", so], stdout=subprocess.PIPE, text=True).stdout
    assert " o_gate" not in out and "o_bootstrap_woks" not in out
    ldd = subprocess.run(["ldd", so], stdout=subprocess.PIPE, text=True).stdout
    assert "liboracle" not in ldd
    for root, _, files in os.walk(os.path.join(ROOT, "ie-ache_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(root, f)).read()
                for needle in ("oracle_bind", "liboracle", "tfhe_oracle.h", "oracle/", "o_gate(", "o_keygen("):
                    assert needle not in text, (f, needle)
