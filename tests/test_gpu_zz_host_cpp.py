"""GPU: the reference's boundary tests (test_fft / check_fft, test_commit, test_homomorphic_poly_com_elem, test_pcs_eval,
apply_blind_factors, the Lagrange SRS) restated in C++ on include/uzkge_host.hpp, every expected value from the CPU oracle."""
import pytest

from host_cpp import run

pytestmark = pytest.mark.gpu


def test_cpp_host_layer_matches_the_oracle(gpu):
    out = run("gpu")
    assert out.returncode == 0 and "PASS gpu" in out.stdout, out.stdout + out.stderr
