"""Builds tests/host/host_api_test: the reference's boundary tests restated in C++ on include/uzkge_host.hpp (the compiled-language
host layer above the C ABI), linked against the CUDA library and -- as the checker -- the CPU oracle."""
import os
import subprocess

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
BIN = os.path.join(ROOT, "tests", "host", "host_api_test")
SRC = os.path.join(ROOT, "tests", "host", "host_api_test.cpp")
HDRS = [os.path.join(ROOT, "include", h) for h in ("uzkge_host.hpp", "uzkge_transcript.hpp", "uzkge_cuda.h")]


def build() -> str:
    from oracle import cpu
    from uzkge_b200 import ffi

    from uzkge_b200.transcript import host_lib

    cpu.lib()          # builds oracle/liboracle.so when missing
    ffi.lib()          # fails loudly when the CUDA library is not built
    host_lib()         # libuzkge_host.so: Keccak-256, ChaCha20
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < newest:
        cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", BIN,
               "-L", os.path.join(ROOT, "uzkge_b200", "lib"), "-luzkge_cuda", "-luzkge_host", "-L", os.path.join(ROOT, "oracle"), "-l:liboracle.so",
               "-Wl,-rpath,$ORIGIN/../../uzkge_b200/lib", "-Wl,-rpath,$ORIGIN/../../oracle"]
        subprocess.run(cmd, check=True, cwd=ROOT)
    return BIN


def run(mode: str, *args: str) -> subprocess.CompletedProcess:
    return subprocess.run([build(), mode, *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
