"""CPU: the C++ host layer (include/uzkge_host.hpp) compiles warning-free against the C ABI, its host-side logic (trimming, negation,
domain-size rules, DegreeError / ParameterError before any device call) holds, and without a GPU every device-backed call fails
loudly with the UzkgeError variant INTEGRATION.md maps it to -- there is no CPU path behind it."""
import pytest

from host_cpp import run


def test_cpp_host_layer_builds_and_refuses_to_run_without_a_device():
    from uzkge_b200 import ffi

    if ffi.lib().uzkge_cuda_device_count() > 0:
        pytest.skip("a CUDA device is visible: the parity run is tests/test_gpu_zz_host_cpp.py")
    out = run("nodevice")
    assert out.returncode == 0 and "PASS nodevice" in out.stdout, out.stdout + out.stderr


def test_cpp_transcript_rng_and_scalar_arithmetic(domain_kat):
    """include/uzkge_transcript.hpp: Keccak known answer, the host Montgomery arithmetic against the oracle's, choose_ks at seed 0
    against the k[1..4] of the reference's verifier keys (golden), and a transcript script against the Python mirror's challenges."""
    from uzkge_b200.rng import FR_MODULUS
    from uzkge_b200.transcript import Transcript

    k = [int(v, 0) for v in domain_kat["52"]["k"]]
    out = run("serial", *[hex(v) for v in k[1:]])
    assert out.returncode == 0 and "PASS serial" in out.stdout, out.stdout + out.stderr
    got = dict(line.split() for line in out.stdout.splitlines() if line.startswith("challenge"))
    tr = Transcript(b"Plonk shuffle Proof")
    tr.append_u64(52)
    tr.append_challenge(k[1])
    tr.append_message((1).to_bytes(32, "big") + (2).to_bytes(32, "big"))          # the G1 generator's transcript bytes
    tr.append_message(bytes(64))                                                   # the identity's
    c1 = tr.get_challenge_field_elem()
    tr.append_single_byte(0x01)
    tr.append_challenge(c1 * k[2] % FR_MODULUS)
    c2 = tr.get_challenge_field_elem()
    assert int(got["challenge1"], 16) == c1 and int(got["challenge2"], 16) == c2
