"""Host check of the arena sizing (gmix_b200/csrc/layout.h): the PPMd heap window MakeLayout picks must keep the text
area and the two unit areas disjoint for every split of the unit budget, also above 16 MiB where the 2000 MiB heap end
is no longer a multiple of the window size."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_ppmd_window_areas_never_alias(tmp_path):
    exe = str(tmp_path / "layout_check")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(HERE, "native", "layout_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "bad 0" in r.stdout
    assert "windows above 16 MiB checked: 0" not in r.stdout
