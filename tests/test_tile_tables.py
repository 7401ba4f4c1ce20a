"""Host logic of the tile engine: the table builders of libkriging_b200/csrc/tile_tables.hpp (the orders in which the
persistent / one-CTA-per-tile launches of Cholesky, LAUUM and the LOO product walk their tiles) are compiled for the CPU
and checked against brute-force tile sets -- every tile exactly once, for 1 .. 313 panels and several block / band sizes."""
import json
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_tile_tables_cover_their_tile_sets(tmp_path):
    exe = str(tmp_path / "tile_tables_selftest")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "cpp", "tile_tables_selftest.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and res["ok"], out.stdout[-2000:]
    assert res["tables_checked"] > 1000
