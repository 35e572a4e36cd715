"""A few-line SPIR-V assembler used only to unit-test oracle/spv_exec.py's control flow
(tests/test_spv_pin.py::test_interpreter_control_flow_and_masks)."""
import struct


def _ins(op, *words):
    return [((len(words) + 1) << 16) | op, *words]


def _str(s):
    b = s.encode() + b"\0"
    b += b"\0" * (-len(b) % 4)
    return list(struct.unpack("<%dI" % (len(b) // 4), b))


def build_loop_module() -> bytes:
    """layout(local_size_x = 8) in; buffer { uint out_[]; };
    void main() { uint g = gl_GlobalInvocationID.x; uint acc = 0;
                  for (uint i = 0; i < g % 5; i++) acc += i * 3;      // divergent trip count
                  if (g % 2 == 0) acc += 1000;                         // divergent selection
                  out_[g] = acc; }"""
    (VOID, FN, UINT, V3, PIN3, GID, PIN1, RT, ST, PST, BUF, PSB, PFU, BOOL, INT,
     C0, C1, C2, C3, C5, C1000, CI0, MAIN, L0, ACC, I, LH, LC, LB, LCONT, LM, LT, LE) = range(1, 34)
    nid = [34]

    def t():
        nid[0] += 1
        return nid[0] - 1

    w = []
    w += _ins(17, 1)                                   # Capability Shader
    w += _ins(14, 0, 1)                                # MemoryModel Logical GLSL450
    w += _ins(15, 5, MAIN, *_str("main"), GID)         # EntryPoint GLCompute
    w += _ins(16, MAIN, 17, 8, 1, 1)                   # LocalSize 8 1 1
    w += _ins(71, GID, 11, 28)                         # BuiltIn GlobalInvocationId
    w += _ins(71, RT, 6, 4)                            # ArrayStride
    w += _ins(72, ST, 0, 35, 0)                        # Offset 0
    w += _ins(71, ST, 3)                               # BufferBlock
    w += _ins(71, BUF, 34, 0) + _ins(71, BUF, 33, 0)   # set 0 binding 0
    w += _ins(19, VOID) + _ins(33, FN, VOID) + _ins(21, UINT, 32, 0) + _ins(23, V3, UINT, 3)
    w += _ins(32, PIN3, 1, V3) + _ins(59, PIN3, GID, 1) + _ins(32, PIN1, 1, UINT)
    w += _ins(29, RT, UINT) + _ins(30, ST, RT) + _ins(32, PST, 2, ST) + _ins(59, PST, BUF, 2)
    w += _ins(32, PSB, 2, UINT) + _ins(32, PFU, 7, UINT) + _ins(20, BOOL) + _ins(21, INT, 32, 1)
    for cid, val in ((C0, 0), (C1, 1), (C2, 2), (C3, 3), (C5, 5), (C1000, 1000)):
        w += _ins(43, UINT, cid, val)
    w += _ins(43, INT, CI0, 0)
    w += _ins(54, VOID, MAIN, 0, FN) + _ins(248, L0)
    w += _ins(59, PFU, ACC, 7) + _ins(59, PFU, I, 7)
    pg, g, n = t(), t(), t()
    w += _ins(65, PIN1, pg, GID, C0) + _ins(61, UINT, g, pg) + _ins(137, UINT, n, g, C5)
    w += _ins(62, ACC, C0) + _ins(62, I, C0) + _ins(249, LH)
    w += _ins(248, LH) + _ins(246, LM, LCONT, 0) + _ins(249, LC)
    li, c = t(), t()
    w += _ins(248, LC) + _ins(61, UINT, li, I) + _ins(176, BOOL, c, li, n) + _ins(250, c, LB, LM)
    li2, m3, a0, a1 = t(), t(), t(), t()
    w += _ins(248, LB) + _ins(61, UINT, li2, I) + _ins(132, UINT, m3, li2, C3) + _ins(61, UINT, a0, ACC)
    w += _ins(128, UINT, a1, a0, m3) + _ins(62, ACC, a1) + _ins(249, LCONT)
    li3, i1 = t(), t()
    w += _ins(248, LCONT) + _ins(61, UINT, li3, I) + _ins(128, UINT, i1, li3, C1) + _ins(62, I, i1) + _ins(249, LH)
    par, ev = t(), t()
    w += _ins(248, LM) + _ins(137, UINT, par, g, C2) + _ins(170, BOOL, ev, par, C0)
    w += _ins(247, LE, 0) + _ins(250, ev, LT, LE)
    a2, a3 = t(), t()
    w += _ins(248, LT) + _ins(61, UINT, a2, ACC) + _ins(128, UINT, a3, a2, C1000) + _ins(62, ACC, a3) + _ins(249, LE)
    a4, po = t(), t()
    w += _ins(248, LE) + _ins(61, UINT, a4, ACC) + _ins(65, PSB, po, BUF, CI0, g) + _ins(62, po, a4)
    w += _ins(253) + _ins(56)
    header = [0x07230203, 0x00010000, 0, nid[0], 0]
    return struct.pack("<%dI" % (len(header) + len(w)), *header, *w)
