"""Host logic of the C modmap driver (modimizer_b200/csrc/shim/modmap_gpu.c) on CPU: the same source, linked against an
oracle-backed stub of the ABI calls it makes (tests/host/modgpu_stub.c - test infrastructure), must print what the STOCK
modmap prints - Q lines, -v seed lines, the colinear-block M lines, their interleaving - and write the same .ref bytes.
The GPU run of the real driver is tests/test_gpu_cli.py."""
import os

import pytest

import harness as H

CHECK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "modmap_hostcheck")


def test_modmap_driver_host_logic(tmp_path):
    stock = H.ref_cli("modmap")
    if not stock or not os.path.exists(CHECK):
        pytest.skip("stock modmap / modmap_hostcheck not built (no /root/reference in the build container)")
    H.modmap_case(str(tmp_path))
    H.modmap_driver_vs_stock(stock, CHECK, str(tmp_path), check_mod=False)
