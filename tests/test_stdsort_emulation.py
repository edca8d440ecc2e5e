"""CPU: the std::sort replay used by the scipy-order CSC mode (csrc/stdsort_emul.h) against
the real libstdc++ std::sort and heap algorithms, on (key, id) pairs with many duplicates."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emulation_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / 'stdsort_check')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-o', exe,
                           os.path.join(ROOT, 'oracle', 'stdsort_check.cpp')])
    out = subprocess.run([exe, '1500'], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith('OK')
