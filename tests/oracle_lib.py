"""Loads the CPU oracle (oracle/liboracle.so) for the tests.  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_api = None


def build_oracle():
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle.cpp", "oracle.h", "Makefile")]
    if (not os.path.exists(ORACLE_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_LIB


def oracle_api():
    global _api
    if _api is None:
        from vkjit_b200._capi import CApi
        _api = CApi(build_oracle(), "orc_")
    return _api


def OracleIr():
    from vkjit_b200.ir import Ir
    return Ir(_api=oracle_api())
