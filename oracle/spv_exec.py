"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A small SPIR-V interpreter that EXECUTES the reference's own shipped shader binaries

    /root/reference/shader/spv/{propagate,fft_row,fft_col,correction}.comp.spv
    /root/reference/shader/spv/{ocean.vert,ocean.frag}.spv

on the CPU, so that the self-authored oracle (ocean_oracle.c / ocean_oracle.py) and the CUDA path
are pinned by something the reference SHIPS AND RUNS, not only by a reading of its GLSL sources.
The reference embeds exactly these binaries (`include_bytes!`, src/fft.rs:20-25, src/ocean.rs:26-28,
195-197) and dispatches them per frame at src/render.rs:1122-1287; `run_reference_frame` below replays
that dispatch sequence with the bindings of src/render.rs:944-988.

Execution model: SIMT lock-step. Every invocation of a dispatch is one lane of a numpy array; an
instruction is executed for all active lanes at once; structured control flow (OpSelectionMerge /
OpLoopMerge as glslang emits them) is followed with per-lane activity masks; OpControlBarrier is a
no-op because lock-step execution of a data-race-free program is one of its valid interleavings.
`Workgroup` variables get one copy per workgroup, `Function` variables one per lane.

Arithmetic: every floating-point instruction is evaluated in IEEE binary32, one rounding per SPIR-V
instruction (no FMA contraction: the modules carry no contraction-enabling decorations a driver would
need, and un-contracted evaluation is always a valid one). GLSL.std.450 Sin/Cos/Pow are evaluated in
binary64 and rounded once to binary32, i.e. "correctly rounded" -- inside every Vulkan driver's
allowed error (sin/cos: 2^-11 absolute), which is the only latitude the reference's results have.
Integer instructions wrap mod 2^32 (`2*gid - resolution - 1`, propagate.comp:45-46).

Only what the six reference modules use is implemented (about 60 opcodes); anything else raises.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------
# module parsing
# --------------------------------------------------------------------------------------

OP = {
    3: "Source", 4: "SourceExtension", 5: "Name", 6: "MemberName", 11: "ExtInstImport", 12: "ExtInst",
    14: "MemoryModel", 15: "EntryPoint", 16: "ExecutionMode", 17: "Capability",
    19: "TypeVoid", 20: "TypeBool", 21: "TypeInt", 22: "TypeFloat", 23: "TypeVector", 24: "TypeMatrix",
    25: "TypeImage", 26: "TypeSampler", 27: "TypeSampledImage", 28: "TypeArray", 29: "TypeRuntimeArray",
    30: "TypeStruct", 32: "TypePointer", 33: "TypeFunction",
    41: "ConstantTrue", 42: "ConstantFalse", 43: "Constant", 44: "ConstantComposite",
    54: "Function", 55: "FunctionParameter", 56: "FunctionEnd", 57: "FunctionCall",
    59: "Variable", 61: "Load", 62: "Store", 65: "AccessChain", 71: "Decorate", 72: "MemberDecorate",
    79: "VectorShuffle", 80: "CompositeConstruct", 81: "CompositeExtract",
    86: "SampledImage", 87: "ImageSampleImplicitLod", 88: "ImageSampleExplicitLod", 99: "ImageWrite",
    111: "ConvertSToF", 112: "ConvertUToF", 124: "Bitcast",
    127: "FNegate", 128: "IAdd", 129: "FAdd", 130: "ISub", 131: "FSub", 132: "IMul", 133: "FMul",
    136: "FDiv", 137: "UMod", 142: "VectorTimesScalar", 145: "MatrixTimesVector", 146: "MatrixTimesMatrix",
    148: "Dot", 169: "Select", 170: "IEqual", 176: "ULessThan", 186: "FOrdGreaterThan",
    196: "ShiftLeftLogical", 199: "BitwiseAnd", 224: "ControlBarrier", 225: "MemoryBarrier",
    246: "LoopMerge", 247: "SelectionMerge", 248: "Label", 249: "Branch", 250: "BranchConditional",
    253: "Return", 254: "ReturnValue",
}

# storage classes
SC_UNIFORM_CONSTANT, SC_INPUT, SC_UNIFORM, SC_OUTPUT, SC_WORKGROUP, SC_PRIVATE, SC_FUNCTION = 0, 1, 2, 3, 4, 6, 7
# decorations
DEC_BUILTIN, DEC_LOCATION, DEC_BINDING, DEC_DESCRIPTOR_SET = 11, 30, 33, 34
BUILTIN_POSITION, BUILTIN_GLOBAL_INVOCATION_ID = 0, 28


@dataclass
class Type:
    kind: str                      # void bool int float vec mat array rtarray struct ptr image sampler sampledimage func
    elem: "Type | None" = None
    count: int = 0
    signed: bool = False
    members: list = field(default_factory=list)
    storage: int = -1

    def shape(self):
        """numpy trailing shape of one value of this type."""
        if self.kind in ("bool", "int", "float"):
            return ()
        if self.kind in ("vec", "array", "mat"):
            return (self.count,) + self.elem.shape()
        raise TypeError(f"no array shape for type {self.kind}")

    def dtype(self):
        if self.kind == "bool":
            return np.bool_
        if self.kind == "int":
            return np.int32 if self.signed else np.uint32
        if self.kind == "float":
            return np.float32
        if self.kind in ("vec", "array", "mat"):
            return self.elem.dtype()
        raise TypeError(f"no dtype for type {self.kind}")


@dataclass
class Inst:
    op: str
    rtype: int          # result type id (0 if none)
    rid: int            # result id (0 if none)
    args: tuple


@dataclass
class Function:
    fid: int
    rtype: int
    params: list
    blocks: dict        # label id -> list[Inst]
    entry: int
    variables: list     # OpVariable instructions of storage class Function


_HAS_TYPE_AND_RESULT = {
    "ExtInst", "ConstantTrue", "ConstantFalse", "Constant", "ConstantComposite", "Function", "FunctionParameter",
    "FunctionCall", "Variable", "Load", "AccessChain", "VectorShuffle", "CompositeConstruct", "CompositeExtract",
    "SampledImage", "ImageSampleImplicitLod", "ImageSampleExplicitLod", "ConvertSToF", "ConvertUToF", "Bitcast",
    "FNegate", "IAdd", "FAdd", "ISub", "FSub", "IMul", "FMul", "FDiv", "UMod", "VectorTimesScalar",
    "MatrixTimesVector", "MatrixTimesMatrix", "Dot", "Select", "IEqual", "ULessThan", "FOrdGreaterThan",
    "ShiftLeftLogical", "BitwiseAnd",
}
_HAS_RESULT_ONLY = {"ExtInstImport", "TypeVoid", "TypeBool", "TypeInt", "TypeFloat", "TypeVector", "TypeMatrix",
                    "TypeImage", "TypeSampler", "TypeSampledImage", "TypeArray", "TypeRuntimeArray", "TypeStruct",
                    "TypePointer", "TypeFunction", "Label"}


def _string(words):
    raw = b"".join(struct.pack("<I", w) for w in words)
    return raw.split(b"\0")[0].decode()


class Module:
    """A parsed SPIR-V module (logical layout of the binary, spec section 2.3)."""

    def __init__(self, blob: bytes):
        if len(blob) % 4 or len(blob) < 20:
            raise ValueError("not a SPIR-V binary")
        w = struct.unpack("<%dI" % (len(blob) // 4), blob)
        if w[0] != 0x07230203:
            raise ValueError("bad SPIR-V magic")
        self.version, self.generator, self.bound = w[1], w[2], w[3]
        self.names, self.member_names = {}, {}
        self.decorations = {}          # id -> {decoration: operands}
        self.types, self.constants = {}, {}
        self.global_vars = {}          # id -> (type id, storage class)
        self.functions = {}
        self.entry_point = None
        self.execution_model = None
        self.local_size = (1, 1, 1)
        self.opcode_census = {}
        cur, cur_label = None, None
        i = 5
        while i < len(w):
            wc, opc = w[i] >> 16, w[i] & 0xFFFF
            if wc == 0:
                raise ValueError("zero word count")
            a = w[i + 1:i + wc]
            i += wc
            if opc not in OP:
                raise NotImplementedError(f"SPIR-V opcode {opc} is not implemented")
            name = OP[opc]
            self.opcode_census[name] = self.opcode_census.get(name, 0) + 1
            if name in _HAS_TYPE_AND_RESULT:
                ins = Inst(name, a[0], a[1], tuple(a[2:]))
            elif name in _HAS_RESULT_ONLY:
                ins = Inst(name, 0, a[0], tuple(a[1:]))
            else:
                ins = Inst(name, 0, 0, tuple(a))
            if name in ("Capability", "MemoryModel", "Source", "SourceExtension", "ExtInstImport", "MemberName"):
                if name == "Capability" and a[0] != 1:
                    raise NotImplementedError(f"capability {a[0]}")
                continue
            if name == "Name":
                self.names[a[0]] = _string(a[1:])
            elif name == "EntryPoint":
                self.execution_model, self.entry_point = a[0], a[1]
            elif name == "ExecutionMode":
                if a[1] == 17:          # LocalSize
                    self.local_size = (a[2], a[3], a[4])
            elif name == "Decorate":
                self.decorations.setdefault(a[0], {})[a[1]] = tuple(a[2:])
            elif name == "MemberDecorate":
                pass
            elif name.startswith("Type"):
                self._add_type(ins)
            elif name in ("Constant", "ConstantTrue", "ConstantFalse", "ConstantComposite"):
                self._add_constant(ins)
            elif name == "Variable" and cur is None:
                self.global_vars[ins.rid] = (ins.rtype, ins.args[0])
            elif name == "Function":
                cur = Function(ins.rid, ins.rtype, [], {}, 0, [])
            elif name == "FunctionParameter":
                cur.params.append(ins.rid)
            elif name == "FunctionEnd":
                self.functions[cur.fid] = cur
                cur, cur_label = None, None
            elif name == "Label":
                cur_label = ins.rid
                cur.blocks[cur_label] = []
                if not cur.entry:
                    cur.entry = cur_label
            elif cur is not None:
                if name == "Variable":
                    cur.variables.append(ins)
                else:
                    cur.blocks[cur_label].append(ins)
            else:
                raise NotImplementedError(f"{name} at module scope")

    def _add_type(self, ins):
        k, a, t = ins.op, ins.args, self.types
        if k == "TypeVoid":
            ty = Type("void")
        elif k == "TypeBool":
            ty = Type("bool")
        elif k == "TypeInt":
            if a[0] != 32:
                raise NotImplementedError("only 32-bit integers")
            ty = Type("int", signed=bool(a[1]))
        elif k == "TypeFloat":
            if a[0] != 32:
                raise NotImplementedError("only 32-bit floats")
            ty = Type("float")
        elif k == "TypeVector":
            ty = Type("vec", elem=t[a[0]], count=a[1])
        elif k == "TypeMatrix":
            ty = Type("mat", elem=t[a[0]], count=a[1])      # [column][row]
        elif k == "TypeArray":
            ty = Type("array", elem=t[a[0]], count=int(self.constants[a[1]][0]))
        elif k == "TypeRuntimeArray":
            ty = Type("rtarray", elem=t[a[0]])
        elif k == "TypeStruct":
            ty = Type("struct", members=[t[m] for m in a])
        elif k == "TypePointer":
            ty = Type("ptr", elem=t[a[1]], storage=a[0])
        elif k == "TypeImage":
            ty = Type("image")
        elif k == "TypeSampler":
            ty = Type("sampler")
        elif k == "TypeSampledImage":
            ty = Type("sampledimage")
        elif k == "TypeFunction":
            ty = Type("func")
        else:
            raise NotImplementedError(k)
        t[ins.rid] = ty

    def _add_constant(self, ins):
        ty = self.types[ins.rtype]
        if ins.op == "ConstantTrue":
            v = np.array([True])
        elif ins.op == "ConstantFalse":
            v = np.array([False])
        elif ins.op == "Constant":
            v = np.array([ins.args[0]], dtype=np.uint32).view(ty.dtype())
        else:
            v = np.stack([self.constants[c][0] for c in ins.args])[None]
        self.constants[ins.rid] = v          # leading lane axis of length 1 (broadcasts)

    def find_var(self, name):
        for vid in self.global_vars:
            if self.names.get(vid) == name:
                return vid
        raise KeyError(name)


def load_module(path: str) -> Module:
    with open(path, "rb") as f:
        return Module(f.read())


# --------------------------------------------------------------------------------------
# resources
# --------------------------------------------------------------------------------------

class Image2D:
    """A 2-D RGBA32F image, texel (x, y) at data[y, x]. Sampling follows the Vulkan texel-filtering
    equations for VK_FILTER_LINEAR + REPEAT addressing -- what the reference's sampler is
    (`SamplerDesc::new(Filter::Linear, WrapMode::Tile)`, src/render.rs:397-398) -- in binary32."""

    def __init__(self, data):
        self.data = data

    def sample_linear_tile(self, uv, offset=(0, 0)):
        h, w = self.data.shape[:2]
        f = np.float32
        u = uv[:, 0] * f(w) - f(0.5)
        v = uv[:, 1] * f(h) - f(0.5)
        i0 = np.floor(u)
        j0 = np.floor(v)
        a = (u - i0).astype(f)[:, None]
        b = (v - j0).astype(f)[:, None]
        i0 = i0.astype(np.int64) + int(offset[0])
        j0 = j0.astype(np.int64) + int(offset[1])
        i1, j1 = (i0 + 1) % w, (j0 + 1) % h
        i0, j0 = i0 % w, j0 % h
        d = self.data
        one = f(1.0)
        top = d[j0, i0] * (one - a) + d[j0, i1] * a
        bot = d[j1, i0] * (one - a) + d[j1, i1] * a
        return (top * (one - b) + bot * b).astype(f)


@dataclass
class Ptr:
    var: int            # variable id (global or function-local)
    idx: tuple          # per-level indices: python ints or (L,) integer arrays


# --------------------------------------------------------------------------------------
# execution
# --------------------------------------------------------------------------------------

class Executor:
    """Runs one entry point for L lanes in lock-step.

    resources: {(set, binding): obj}; obj is
        * a list of numpy arrays, one per struct member (buffer / uniform blocks; a runtime array of vec2 is
          an (n, 2) float32 array, a scalar member a 0-d or (1,) array),
        * an Image2D (storage or sampled image), or any object for a sampler.
    inputs:    {("builtin", id) | ("location", n): (L, ...) array}
    """

    def __init__(self, module: Module, lanes: int, resources: dict, inputs: dict, workgroup_of_lane=None,
                 n_workgroups: int = 1):
        self.m, self.L = module, lanes
        self.lane = np.arange(lanes)
        self.wg = workgroup_of_lane if workgroup_of_lane is not None else np.zeros(lanes, np.int64)
        self.mem = {}            # variable id -> backing store
        self.storage = {}        # variable id -> storage class
        self.outputs = {}
        for vid, (tid, sc) in module.global_vars.items():
            pointee = module.types[tid].elem
            dec = module.decorations.get(vid, {})
            self.storage[vid] = sc
            if sc in (SC_UNIFORM, SC_UNIFORM_CONSTANT):
                key = (dec.get(DEC_DESCRIPTOR_SET, (0,))[0], dec[DEC_BINDING][0])
                if key not in resources:
                    raise KeyError(f"no resource bound at set {key[0]} binding {key[1]} ({module.names.get(vid)})")
                self.mem[vid] = resources[key]
            elif sc == SC_INPUT:
                key = ("builtin", dec[DEC_BUILTIN][0]) if DEC_BUILTIN in dec else ("location", dec[DEC_LOCATION][0])
                self.mem[vid] = np.ascontiguousarray(inputs[key], dtype=pointee.dtype())
            elif sc == SC_OUTPUT:
                if pointee.kind == "struct":      # gl_PerVertex: keep one array per member
                    self.mem[vid] = [np.zeros((lanes,) + mt.shape(), mt.dtype()) for mt in pointee.members]
                else:
                    self.mem[vid] = np.zeros((lanes,) + pointee.shape(), pointee.dtype())
                self.outputs[vid] = self.mem[vid]
            elif sc == SC_WORKGROUP:
                self.mem[vid] = np.zeros((n_workgroups,) + pointee.shape(), pointee.dtype())
            else:
                raise NotImplementedError(f"storage class {sc}")
        self.named_locals = {}   # name -> array of the main function's Function variables after the run

    # ---- memory -------------------------------------------------------------------
    def _resolve(self, p: Ptr):
        """-> (backing array, index tuple) with the lane / workgroup axis made explicit."""
        sc = self.storage[p.var]
        store, idx = self.mem[p.var], p.idx
        if isinstance(store, list):                      # struct of member arrays: first index is constant
            store, idx = store[int(idx[0])], idx[1:]
            if sc in (SC_UNIFORM, SC_UNIFORM_CONSTANT):
                return store, idx, False
            return store, (self.lane,) + idx, True
        if sc in (SC_FUNCTION, SC_INPUT, SC_OUTPUT, SC_PRIVATE):
            return store, (self.lane,) + idx, True
        if sc == SC_UNIFORM_CONSTANT:
            return store, idx, False                     # image / sampler handle
        if sc == SC_WORKGROUP:
            return store, (self.wg,) + idx, True
        raise NotImplementedError(f"pointer into storage class {sc}")

    def load(self, p: Ptr):
        store, idx, per_lane = self._resolve(p)
        if not isinstance(store, np.ndarray):
            return store                                 # image / sampler handle
        if per_lane:
            return store[idx]
        if not idx:
            return np.asarray(store)[None]               # whole member, broadcast over lanes
        if all(isinstance(i, (int, np.integer)) for i in idx):
            return np.asarray(store[idx])[None]
        return store[idx]

    def store(self, p: Ptr, value, mask):
        store, idx, per_lane = self._resolve(p)
        if not per_lane:
            idx = tuple(np.broadcast_to(i, (self.L,)) if isinstance(i, np.ndarray) else np.full(self.L, i) for i in idx)
        value = np.broadcast_to(value, (self.L,) + np.shape(value)[1:])
        if mask is None:
            store[idx] = value
        else:
            sel = tuple(i[mask] if isinstance(i, np.ndarray) else i for i in idx)
            store[sel] = value[mask]

    # ---- values -------------------------------------------------------------------
    def val(self, env, i):
        if i in env:
            return env[i]
        if i in self.m.constants:
            return self.m.constants[i]
        if i in self.m.global_vars:
            return Ptr(i, ())
        raise KeyError(f"%{i} has no value")

    @staticmethod
    def _u(x):
        x = np.asarray(x)
        return x.view(np.uint32) if x.dtype == np.int32 else x.astype(np.uint32, copy=False)

    def _as(self, tid, x):
        dt = self.m.types[tid].dtype()
        x = np.asarray(x)
        if x.dtype == dt:
            return x
        if x.dtype.kind in "iu" and np.dtype(dt).kind in "iu":
            return x.view(dt)
        return x.astype(dt)

    # ---- function / block execution --------------------------------------------------
    def run(self):
        f = self.m.functions[self.m.entry_point]
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            self.call(f, [], None, top=True)
        return self

    def call(self, f: Function, args, mask, top=False):
        env = dict(zip(f.params, args))
        for v in f.variables:
            pointee = self.m.types[v.rtype].elem
            self.mem[v.rid] = np.zeros((self.L,) + pointee.shape(), pointee.dtype())
            self.storage[v.rid] = SC_FUNCTION
            env[v.rid] = Ptr(v.rid, ())
        ret = self.flow(f, env, f.entry, mask, None)
        if top:
            for v in f.variables:
                nm = self.m.names.get(v.rid)
                if nm:
                    self.named_locals[nm] = self.mem[v.rid]
        return ret

    def flow(self, f, env, label, mask, stop):
        """Execute blocks from `label` until control reaches `stop` (exclusive) or the function returns."""
        while label != stop:
            block = f.blocks[label]
            merge = None
            for ins in block:
                op = ins.op
                if op == "SelectionMerge":
                    merge = ("sel", ins.args[0])
                elif op == "LoopMerge":
                    merge = ("loop", ins.args[0], ins.args[1])
                elif op == "Branch":
                    if merge and merge[0] == "loop":
                        self.loop(f, env, header=label, first=ins.args[0], merge=merge[1], cont=merge[2], mask=mask)
                        label = merge[1]
                    else:
                        label = ins.args[0]
                    break
                elif op == "BranchConditional":
                    if not merge or merge[0] != "sel":
                        raise NotImplementedError("conditional branch without a selection merge")
                    c = np.broadcast_to(self.val(env, ins.args[0]), (self.L,))
                    base = np.ones(self.L, bool) if mask is None else mask
                    for tgt, m_ in ((ins.args[1], base & c), (ins.args[2], base & ~c)):
                        if tgt != merge[1] and m_.any():
                            if self.flow(f, env, tgt, m_, merge[1]) is not None:
                                raise NotImplementedError("return inside a selection")
                    label = merge[1]
                    break
                elif op == "Return":
                    if stop is not None:
                        raise NotImplementedError("return inside a structured construct")
                    return None
                elif op == "ReturnValue":
                    if stop is not None:
                        raise NotImplementedError("return inside a structured construct")
                    return self.val(env, ins.args[0])
                else:
                    self.exec(ins, env, mask)
            else:
                raise ValueError("block without terminator")
        return None

    def loop(self, f, env, header, first, merge, cont, mask):
        """glslang loop shape: header -> cond block (BranchConditional body/merge) -> body -> cont -> header."""
        active = np.ones(self.L, bool) if mask is None else mask.copy()
        guard = 0
        while True:
            guard += 1
            if guard > 1 << 20:
                raise RuntimeError("loop does not terminate")
            label = first
            # walk to the loop's exit test
            while True:
                block = f.blocks[label]
                term = block[-1]
                for ins in block[:-1]:
                    if ins.op in ("SelectionMerge", "LoopMerge"):
                        raise NotImplementedError("nested construct before the loop condition")
                    self.exec(ins, env, active)
                if term.op == "BranchConditional":
                    break
                if term.op != "Branch":
                    raise NotImplementedError("unexpected terminator in loop header")
                label = term.args[0]
            c = np.broadcast_to(self.val(env, term.args[0]), (self.L,))
            if term.args[2] == merge:
                body, stay = term.args[1], c
            elif term.args[1] == merge:
                body, stay = term.args[2], ~c
            else:
                raise NotImplementedError("loop condition does not exit to the merge block")
            active = active & stay
            if not active.any():
                return
            self.flow(f, env, body, active, cont)
            self.flow(f, env, cont, active, header)

    # ---- one instruction -------------------------------------------------------------
    def exec(self, ins: Inst, env, mask):
        op, a, f32 = ins.op, ins.args, np.float32
        v = lambda k: self.val(env, a[k])          # noqa: E731
        r = None
        if op in ("MemoryBarrier", "ControlBarrier"):
            return                                   # lock-step: every lane is already here
        if op == "AccessChain":
            base = v(0)
            idx = []
            for k in range(1, len(a)):
                x = self.val(env, a[k])
                x = np.asarray(x)
                if x.shape[0] == 1:
                    idx.append(int(x[0]))
                else:
                    idx.append(self._u(x).astype(np.int64))
            r = Ptr(base.var, base.idx + tuple(idx))
        elif op == "Load":
            r = self.load(v(0))
        elif op == "Store":
            self.store(v(0), v(1), mask)
            return
        elif op == "FunctionCall":
            r = self.call(self.m.functions[a[0]], [self.val(env, x) for x in a[1:]], mask)
            if r is None:
                return
        elif op == "Bitcast":
            r = self._as(ins.rtype, v(0))
        elif op == "ConvertUToF":
            r = self._u(v(0)).astype(f32)
        elif op == "ConvertSToF":
            r = np.asarray(v(0)).view(np.int32).astype(f32)
        elif op in ("IAdd", "ISub", "IMul", "ShiftLeftLogical", "BitwiseAnd", "UMod"):
            x, y = self._u(v(0)), self._u(v(1))
            if op == "IAdd":
                r = x + y
            elif op == "ISub":
                r = x - y
            elif op == "IMul":
                r = x * y
            elif op == "ShiftLeftLogical":
                r = np.left_shift(x, y & np.uint32(31))
            elif op == "BitwiseAnd":
                r = x & y
            else:
                r = x % y
            r = self._as(ins.rtype, r.astype(np.uint32))
        elif op in ("FAdd", "FSub", "FMul", "FDiv"):
            x, y = v(0), v(1)
            assert x.dtype == f32 and y.dtype == f32
            r = {"FAdd": np.add, "FSub": np.subtract, "FMul": np.multiply, "FDiv": np.divide}[op](x, y)
        elif op == "FNegate":
            r = -v(0)
        elif op == "VectorTimesScalar":
            r = v(0) * v(1)[:, None]
        elif op == "Dot":
            x, y = v(0), v(1)
            p = x * y
            r = p[:, 0]
            for k in range(1, p.shape[1]):
                r = r + p[:, k]
        elif op == "MatrixTimesVector":              # M[col][row]: result = sum_c M[c] * v[c]
            M, x = v(0), v(1)
            r = M[:, 0] * x[:, 0:1]
            for c in range(1, M.shape[1]):
                r = r + M[:, c] * x[:, c:c + 1]
        elif op == "MatrixTimesMatrix":
            A, B = v(0), v(1)
            cols = []
            for j in range(B.shape[1]):
                col = A[:, 0] * B[:, j, 0:1]
                for c in range(1, A.shape[1]):
                    col = col + A[:, c] * B[:, j, c:c + 1]
                cols.append(col)
            r = np.stack(cols, axis=1)
        elif op == "CompositeConstruct":
            parts = [np.asarray(self.val(env, x)) for x in a]
            L = max(p.shape[0] for p in parts)
            cols = []
            for p in parts:
                p = np.broadcast_to(p, (L,) + p.shape[1:])
                cols.extend([p] if p.ndim == 1 else [p[:, k] for k in range(p.shape[1])])
            r = np.stack(cols, axis=1)
        elif op == "CompositeExtract":
            r = v(0)
            for k in a[1:]:
                r = r[:, k]
        elif op == "VectorShuffle":
            x, y = v(0), v(1)
            L = max(x.shape[0], y.shape[0])
            both = np.concatenate([np.broadcast_to(x, (L, x.shape[1])), np.broadcast_to(y, (L, y.shape[1]))], axis=1)
            r = both[:, list(a[2:])]
        elif op == "Select":
            c, x, y = v(0), v(1), v(2)
            r = np.where(c[:, None] if max(x.ndim, y.ndim) > c.ndim else c, x, y)
        elif op == "IEqual":
            r = self._u(v(0)) == self._u(v(1))
        elif op == "ULessThan":
            r = self._u(v(0)) < self._u(v(1))
        elif op == "FOrdGreaterThan":
            r = v(0) > v(1)
        elif op == "ExtInst":
            r = self.glsl450(a[1], [self.val(env, x) for x in a[2:]])
        elif op == "SampledImage":
            r = v(0)                                  # the image; the sampler is fixed (linear / tile)
        elif op in ("ImageSampleImplicitLod", "ImageSampleExplicitLod"):
            img, uv = v(0), v(1)
            operands, rest = (a[2], a[3:]) if len(a) > 2 else (0, ())
            off = (0, 0)
            k = 0
            if operands & 2:                          # Lod: single-level image, ignored
                k += 1
            if operands & 8:                          # ConstOffset
                off = tuple(int(t) for t in self.m.constants[rest[k]][0])
                k += 1
            if operands & ~(2 | 8):
                raise NotImplementedError(f"image operands {operands}")
            r = img.sample_linear_tile(np.broadcast_to(uv, (self.L, 2)), off)
        elif op == "ImageWrite":
            img, xy, texel = v(0), v(1), v(2)
            xy = np.broadcast_to(xy, (self.L, 2)).astype(np.int64)
            texel = np.broadcast_to(texel, (self.L, 4))
            if mask is None:
                img.data[xy[:, 1], xy[:, 0]] = texel
            else:
                img.data[xy[mask, 1], xy[mask, 0]] = texel[mask]
            return
        else:
            raise NotImplementedError(op)
        if isinstance(r, np.ndarray) and ins.rtype and self.m.types[ins.rtype].kind in ("float", "vec", "mat"):
            if self.m.types[ins.rtype].dtype() == np.float32 and r.dtype != np.float32:
                raise AssertionError(f"{op} produced {r.dtype}")
        env[ins.rid] = r

    @staticmethod
    def _length(x):
        s = x[:, 0] * x[:, 0]
        for k in range(1, x.shape[1]):
            s = s + x[:, k] * x[:, k]
        return np.sqrt(s)

    def glsl450(self, n, x):
        f32, f64 = np.float32, np.float64
        if n == 13:
            return np.sin(x[0].astype(f64)).astype(f32)
        if n == 14:
            return np.cos(x[0].astype(f64)).astype(f32)
        if n == 26:
            return np.power(x[0].astype(f64), x[1].astype(f64)).astype(f32)
        if n == 40:
            return np.maximum(x[0], x[1])
        if n == 43:
            return np.minimum(np.maximum(x[0], x[1]), x[2])
        if n == 46:                                   # FMix: x*(1-a) + y*a
            return x[0] * (f32(1.0) - x[2]) + x[1] * x[2]
        if n == 66:
            return self._length(x[0])
        if n == 68:
            p, q = x[0], x[1]
            return np.stack([p[:, 1] * q[:, 2] - q[:, 1] * p[:, 2], p[:, 2] * q[:, 0] - q[:, 2] * p[:, 0],
                             p[:, 0] * q[:, 1] - q[:, 0] * p[:, 1]], axis=1)
        if n == 69:
            return x[0] / self._length(x[0])[:, None]
        raise NotImplementedError(f"GLSL.std.450 instruction {n}")


# --------------------------------------------------------------------------------------
# dispatch helpers
# --------------------------------------------------------------------------------------

def dispatch(module: Module, groups, resources):
    """vkCmdDispatch(groups) of a compute module: all invocations of all workgroups in lock-step."""
    if module.execution_model != 5:
        raise ValueError("not a compute module")
    lx, ly, lz = module.local_size
    gx, gy, gz = groups
    nx, ny, nz = gx * lx, gy * ly, gz * lz
    z, y, x = np.meshgrid(np.arange(nz, dtype=np.uint32), np.arange(ny, dtype=np.uint32),
                          np.arange(nx, dtype=np.uint32), indexing="ij")
    gid = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    wg = (gid[:, 0] // lx).astype(np.int64) + gx * ((gid[:, 1] // ly).astype(np.int64) + gy * (gid[:, 2] // lz).astype(np.int64))
    ex = Executor(module, gid.shape[0], resources, {("builtin", BUILTIN_GLOBAL_INVOCATION_ID): gid},
                  workgroup_of_lane=wg, n_workgroups=gx * gy * gz)
    return ex.run()


RESOLUTION = 512            # src/render.rs:42-44 (WORKGROUP_SIZE 16 x WORKGROUP_NUM 32)
WORKGROUP_NUM = 32
DOMAIN_SIZE = 1000.0        # src/render.rs:46


def run_reference_frame(spv_dir, spectrum, omega, time, domain_size=DOMAIN_SIZE, keep_spectra=False):
    """The reference's per-frame compute recording (src/render.rs:1101-1310) on its own SPIR-V:
    propagate [32,32,1]; fft_row [1,512,1] x (dx, dy, dz); fft_col [1,512,1] x 3; correction [32,32,1],
    with the descriptor bindings of src/render.rs:944-988. -> displacement image [512, 512, 4] float32
    (and, optionally, the post-propagate spectra (dy=height, dx, dz) as [512*512, 2] arrays)."""
    import os
    n = RESOLUTION
    mods = {k: load_module(os.path.join(spv_dir, k + ".comp.spv")) for k in ("propagate", "fft_row", "fft_col", "correction")}
    initial_spec = np.ascontiguousarray(spectrum, np.float32).reshape(n * n, 2)
    omega_buf = np.ascontiguousarray(omega, np.float32).reshape(n * n)
    dx, dy, dz = (np.zeros((n * n, 2), np.float32) for _ in range(3))
    locals_ = [np.array(time, np.float32), np.array(n, np.int32), np.array(domain_size, np.float32)]   # src/render.rs:1107-1111
    dispatch(mods["propagate"], (WORKGROUP_NUM, WORKGROUP_NUM, 1), {
        (0, 0): locals_, (0, 1): [initial_spec], (0, 2): [omega_buf], (0, 3): [dy], (0, 4): [dx], (0, 5): [dz]})
    spectra = (dy.copy(), dx.copy(), dz.copy()) if keep_spectra else None
    for buf in (dx, dy, dz):                                        # fft.desc_sets[0..3], src/render.rs:1158-1179
        dispatch(mods["fft_row"], (1, n, 1), {(0, 0): [buf]})
    for buf in (dx, dy, dz):                                        # src/render.rs:1210-1231
        dispatch(mods["fft_col"], (1, n, 1), {(0, 0): [buf]})
    image = Image2D(np.zeros((n, n, 4), np.float32))
    dispatch(mods["correction"], (WORKGROUP_NUM, WORKGROUP_NUM, 1), {
        (0, 0): [np.array(n, np.uint32)], (0, 1): [dy], (0, 2): [dx], (0, 3): [dz], (0, 4): image})
    return (image.data, spectra) if keep_spectra else image.data


def run_fragment_normals(spv_dir, displacement, uv, camera_pos=(0.0, 50.0, 0.0), pos_world=None):
    """shader/spv/ocean.frag.spv at the given uv coordinates -> (N, Target0): the shader's local `N`
    (ocean.frag:66) and its colour output. displacement: [H, W, 4] float32 map bound as u_Texture."""
    import os
    m = load_module(os.path.join(spv_dir, "ocean.frag.spv"))
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    L = uv.shape[0]
    pw = np.zeros((L, 3), np.float32) if pos_world is None else np.ascontiguousarray(pos_world, np.float32)
    eye = np.eye(4, dtype=np.float32)
    res = {(0, 0): [eye, eye, np.asarray(camera_pos, np.float32)], (0, 1): Image2D(np.ascontiguousarray(displacement, np.float32)),
           (0, 2): "linear/tile"}
    ex = Executor(m, L, res, {("location", 0): uv, ("location", 1): pw}).run()
    (target,) = ex.outputs.values()
    return ex.named_locals["N"], target


def run_vertex_displacement(spv_dir, displacement, a_pos, a_uv, a_offset):
    """shader/spv/ocean.vert.spv -> p_PosWorld (ocean.vert:21-25,29): a_Pos + sampled displacement
    (y/3, xz/3.5) + patch offset, with identity projection/view."""
    import os
    m = load_module(os.path.join(spv_dir, "ocean.vert.spv"))
    a_pos = np.ascontiguousarray(a_pos, np.float32).reshape(-1, 3)
    L = a_pos.shape[0]
    eye = np.eye(4, dtype=np.float32)
    res = {(0, 0): [eye, eye, np.zeros(3, np.float32)], (0, 1): Image2D(np.ascontiguousarray(displacement, np.float32)),
           (0, 2): "linear/tile"}
    inputs = {("location", 0): a_pos, ("location", 1): np.ascontiguousarray(a_uv, np.float32).reshape(L, 2),
              ("location", 2): np.broadcast_to(np.asarray(a_offset, np.float32), (L, 2))}
    ex = Executor(m, L, res, inputs).run()
    for vid, arr in ex.outputs.items():
        if m.names.get(vid) == "p_PosWorld":
            return arr
    raise KeyError("p_PosWorld")
