"""parameters.in semantics pinned by the REFERENCE's own parser: the same files go through
UpcGenerator::configGeneratorFromFile of the reference (src/UpcGenerator.cpp:183-346, compiled unmodified into oracle/_ref)
and of the drop-in (upcgen_b200/host/UpcGenerator.cpp); the parameter blocks they leave behind -- every member the
parser can touch, printed with 17 digits by ONE function compiled against either set of headers
(oracle/refshim/param_dump.h) -- must be identical, quirks included: comment lines, trailing comments, unknown keys,
repeated keys, blank lines (which re-apply the previous pair), values in scientific notation, SQRTS setting both
Lorentz factors.  CPU only."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HOST = os.path.join(ROOT, "upcgen_b200", "host")
from oracle import pyref  # noqa: E402

FILES = {
    "the repository's own parameters.in, more or less": """NUCLEUS_Z 82
NUCLEUS_A 208
WS_R 6.68
WS_A 0.447
SQRTS 5020
PROC_ID 15
LEP_A 0
NEVENTS 1000
DO_PT_CUT 0
PT_MIN 0
DO_ETA_CUT 0
ETA_MIN -1.0
ETA_MAX 1.0
ZMIN -1
ZMAX 1
MMIN 3.56
MMAX 50
YMIN -6
YMAX 6
BINS_Z 100
BINS_M 1001
BINS_Y 121
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 1
USE_POLARIZED_CS 0
PYTHIA_VERSION 8
PYTHIA8_FSR 1
PYTHIA8_DECAYS 0
SEED 0
USE_ROOT_OUTPUT 1
USE_HEPMC_OUTPUT 0
""",
    "comments, blanks, unknown and repeated keys": """# a comment line
NUCLEUS_Z 54   # xenon
NUCLEUS_A 129 trailing words are ignored
SQRTS 5.44e3

   PROC_ID 51
ALP_MASS 1.5
ALP_WIDTH 1e-2
NOT_A_PARAMETER 17
BINS_M 77
BINS_M 78
#BINS_Y 5
BINS_Y 33
SEED 123456789012
USE_HEPMC_OUTPUT 1
DO_M_CUT 1
LOW_M_CUT 0.3
HIGH_M_CUT 4.5e0
SHADOWING 4
DECAY_PDG 11
""",
    "polarised, cuts on, numbers written as floats where ints are read": """PROC_ID 11
USE_POLARIZED_CS 1
DO_PT_CUT 1
PT_MIN 0.35
DO_ETA_CUT 1
ETA_MIN -2.5
ETA_MAX 2.5
NEVENTS 25
MMIN 0.4
MMAX 12.5
YMIN -4.5
YMAX 4.5
ZMIN -0.9
ZMAX 0.95
BINS_Z 40
FLUX_POINT 0
BREAKUP_MODE 3
NON_ZERO_GAM_PT 0
LEP_A 0.0011
WS_R 5.36
WS_A 0.59
PYTHIA_VERSION 6
PYTHIA8_FSR 0
PYTHIA8_DECAYS 1
""",
}


def _reference_dump(path):
    L = C.CDLL(pyref.SO)
    L.upcrefgen_parse.restype = C.c_long
    L.upcrefgen_parse.argtypes = [C.c_char_p, C.c_char_p, C.c_long]
    buf = C.create_string_buffer(1 << 14)
    n = L.upcrefgen_parse(path.encode(), buf, len(buf))
    assert n > 100
    return buf.value.decode()


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", list(FILES))
def test_parser_equals_the_references_own(name, tmp_path):
    subprocess.check_call(["make", "-s", "-C", HOST])
    exe = os.path.join(ROOT, "tests", "cpp", "parse_check")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I", HOST, "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "parse_check.cpp"), "-L", HOST, "-lupcgen_host",
                           "-L", os.path.join(ROOT, "upcgen_b200"), "-lupcgpu", f"-Wl,-rpath,{HOST}",
                           f"-Wl,-rpath,{os.path.join(ROOT, 'upcgen_b200')}"])
    par = tmp_path / "parameters.in"
    par.write_text(FILES[name])
    mine = subprocess.run([exe, str(par)], capture_output=True, text=True, timeout=60)
    assert mine.returncode == 0, mine.stderr
    # the reference's statics (Z, A, R, a, sqrts, g1, g2) live for the life of the process: one process per file
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests'); from test_reference_params import "
            "_reference_dump; print(_reference_dump(%r), end='')" % (ROOT, ROOT, str(par)))
    ref = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert ref.returncode == 0, ref.stderr[-2000:]
    theirs = "".join(l + "\n" for l in ref.stdout.splitlines() if l and l.split()[0].isupper() and not l.startswith("["))
    ours = mine.stdout
    assert len(ours.splitlines()) >= 40
    assert ours == theirs, "\n".join(f"{a!r} != {b!r}" for a, b in zip(ours.splitlines(), theirs.splitlines()) if a != b)
