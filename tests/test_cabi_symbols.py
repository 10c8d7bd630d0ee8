"""The C-ABI library loads and exports every symbol include/cfpp.h declares (no compute calls: CPU-only check)."""
import ctypes, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'cfpp.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cfpp_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    import __graft_entry__ as g
    g.build()
    from contextflow_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/cfpp.h but not exported by libcfpp.so'
    # and the binding table covers exactly the header
    assert sorted(_cabi._SIGNATURES) == names


def test_binding_loads_and_reports_version():
    from contextflow_b200 import _cabi
    assert _cabi.lib().cfpp_version() >= 100
    assert _cabi.lib().cfpp_vit_layer_floats(52) == 4 * 52 + 52 * 192 + 64 * 64 + 2 * 52 * 64 + 2 * 64
    assert _cabi.launch_count() == 0


def test_struct_layouts_match_header():
    """ctypes mirrors of cfpp_vit_desc / cfpp_enc_desc must have the C layout (checked against a compiled sizeof probe)."""
    import subprocess, tempfile
    from contextflow_b200 import _cabi
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "cfpp.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(cfpp_vit_desc), sizeof(cfpp_enc_desc), offsetof(cfpp_enc_desc, inner_w), offsetof(cfpp_enc_desc, cb3));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 'p.c'), 'w').write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 'p.c'), '-o', os.path.join(d, 'p')])
        out = subprocess.check_output([os.path.join(d, 'p')]).split()
    assert int(out[0]) == ctypes.sizeof(_cabi.VitDesc)
    assert int(out[1]) == ctypes.sizeof(_cabi.EncDesc)
    assert int(out[2]) == _cabi.EncDesc.inner_w.offset
    assert int(out[3]) == _cabi.EncDesc.cb3.offset
