"""CPU: the single-file container `.pipx` (losslessh264_b200/host/pip_container.cpp) — packing a `.pip` file set and
unpacking it again gives back every file byte for byte; truncated and corrupt archives are refused."""
import os

import pytest

from losslessh264_b200 import pip as pipmod
from test_pip_expand import expand_cpu, write_pip_cpu  # noqa: F401  (fixture)


@pytest.mark.parametrize("name", ["tibby.264", "tibbycabac.264", "Error_I_P.264", "syn_qcif_4slices_8f.264"])
def test_pack_unpack_is_lossless(name, tmp_path, expand_cpu):
    d = str(tmp_path)
    src = write_pip_cpu(name, d)
    arc = os.path.join(d, "x.pipx")
    pipmod.pack(src, arc)
    default, tags = pipmod.read_pip(src)
    assert pipmod.read_archive(arc) == (default, tags)                 # the documented layout, parsed independently in Python
    assert os.path.getsize(arc) == 12 + 12 * (1 + len(tags)) + len(default) + sum(len(b) for b in tags.values())
    back = os.path.join(d, "y.pip")
    pipmod.unpack(arc, back)
    assert pipmod.read_pip(back) == (default, tags)
    assert sorted(f[1:] for f in os.listdir(d) if f.startswith("y.pip")) == sorted(f[1:] for f in os.listdir(d) if f.startswith("x.pip") and not f.endswith(".pipx"))


def test_damaged_archives_are_refused(tmp_path, expand_cpu):
    d = str(tmp_path)
    src = write_pip_cpu("tibby.264", d)
    arc = os.path.join(d, "x.pipx")
    pipmod.pack(src, arc)
    b = open(arc, "rb").read()
    for bad in (b[: len(b) // 2], b[:11], b"XXXX" + b[4:], b + b"\0", b[:8] + (1000).to_bytes(4, "little") + b[12:]):
        open(os.path.join(d, "bad.pipx"), "wb").write(bad)
        with pytest.raises(OSError):
            pipmod.unpack(os.path.join(d, "bad.pipx"), os.path.join(d, "z.pip"))
    with pytest.raises(OSError):
        pipmod.unpack(os.path.join(d, "missing.pipx"), os.path.join(d, "z.pip"))
    with pytest.raises(OSError):
        pipmod.pack(os.path.join(d, "missing.pip"), os.path.join(d, "w.pipx"))
