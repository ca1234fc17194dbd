"""CPU checks of the drop-in boundary: the built library loads, exports every symbol that
include/arrow_b200.h declares, fails loudly without a GPU, and the host-side dispatch rules
(no kernels involved) match the reference's."""
import ctypes as C
import os
import re

import pyarrow as pa
import pytest

from arrow_b200 import _cabi as cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "arrow_b200.h")).read()
    return sorted(set(re.findall(r"B2_API\s+[\w\s\*]+?\b(b2_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = header_symbols()
    bound = sorted(name for name, _, _ in cabi.PROTOTYPES)
    assert declared == bound, f"header-only: {set(declared) - set(bound)}; binding-only: {set(bound) - set(declared)}"


def test_library_exports_every_symbol():
    lib = cabi.lib()  # raises if the .so is missing or a symbol is not exported
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.b2_version().startswith(b"arrow_b200")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = cabi.lib()
    h = C.c_void_p()
    st = lib.b2_context_create(0, C.byref(h))
    assert st != 0 and b"no CPU" in lib.b2_last_error()
    from arrow_b200.device import Context, CudaError
    with pytest.raises(CudaError):
        Context(0)


def test_struct_layout_matches_header():
    # B2Array: 3 pointers + 3 int64 + 2 int32 = 56 bytes; B2Scalar 16; options 16
    assert C.sizeof(cabi.B2Array) == 56
    assert C.sizeof(cabi.B2Scalar) == 16
    assert C.sizeof(cabi.B2CastOptions) == 16
    assert C.sizeof(cabi.B2HashAggOptions) == 16
    assert C.sizeof(cabi.B2Value) == 16
    assert C.sizeof(cabi.B2ReduceResult) == 56


def test_struct_layout_against_the_c_compiler(tmp_path):
    """sizeof / offsetof of every struct as gcc sees include/arrow_b200.h vs the ctypes mirror."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"B2Array": cabi.B2Array, "B2Scalar": cabi.B2Scalar, "B2Value": cabi.B2Value, "B2CastOptions": cabi.B2CastOptions,
               "B2HashAggOptions": cabi.B2HashAggOptions, "B2ReduceResult": cabi.B2ReduceResult}
    lines = []
    for name, st in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "arrow_b200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, st in structs.items():
        assert int(got[name]) == C.sizeof(st), name
        for field, _ in st._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(st, field).offset, f"{name}.{field}"


def test_type_ids_follow_arrow():
    # arrow::Type::type values (type_fwd.h:328-460) as exposed by pyarrow
    from arrow_b200.device import type_id
    for t in (pa.bool_(), pa.uint8(), pa.int8(), pa.uint16(), pa.int16(), pa.uint32(), pa.int32(), pa.uint64(),
              pa.int64(), pa.float32(), pa.float64(), pa.string(), pa.binary(), pa.large_string(), pa.large_binary()):
        assert type_id(t) == t.id, str(t)
    assert type_id(pa.timestamp("us")) == cabi.INT64 and type_id(pa.date32()) == cabi.INT32
    assert type_id(pa.dictionary(pa.int32(), pa.string())) == cabi.INT32


def test_common_numeric_matches_reference_dispatch():
    import itertools
    import pyarrow.compute as pc
    from arrow_b200.compute import common_numeric
    from arrow_b200.device import arrow_type, type_id
    types = [pa.int8(), pa.uint8(), pa.int16(), pa.uint16(), pa.int32(), pa.uint32(), pa.int64(), pa.uint64(),
             pa.float32(), pa.float64()]
    for a, b in itertools.product(types, types):
        want = pc.add(pa.array([1], a), pa.array([1], b)).type  # ArithmeticFunction::DispatchBest
        assert arrow_type(common_numeric([type_id(a), type_id(b)])) == want, f"{a},{b}"


def test_plain_c_client_links_and_calls(tmp_path):
    """The boundary is a C ABI: a C11 translation unit including only include/arrow_b200.h links against
    libarrow_b200.so, and the context either comes up (GPU box) or fails loudly (no CPU fallback)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.join(ROOT, "arrow_b200", "lib")
    src = tmp_path / "client.c"
    src.write_text('''
#include <stdio.h>
#include "arrow_b200.h"
int main(void) {
  B2Context* ctx = NULL;
  printf("version %s\\n", b2_version());
  int st = b2_context_create(0, &ctx);
  if (st != B2_OK) { printf("status %d: %s\\n", st, b2_last_error()); return 0; }
  void* p = NULL;
  if (b2_alloc(ctx, 1 << 20, &p) != B2_OK) return 2;
  b2_free(ctx, p);
  b2_context_destroy(ctx);
  printf("context ok\\n");
  return 0;
}
''')
    exe = tmp_path / "client"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir,
                           "-larrow_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "version arrow_b200" in out.stdout
    assert "context ok" in out.stdout or "no CPU" in out.stdout, out.stdout
