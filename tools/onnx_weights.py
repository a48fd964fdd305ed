#!/usr/bin/env python
"""Rebuild the `best_model` state_dict from the shipped ONNX artefact.

`best_model.pth` is absent from the reference mount (.MISSING_LARGE_BLOBS); the same
486,409 weights ship inside `best_model.onnx` (opset 11, exported from
`src/model.py:LightweightUNet` by `scripts/export_to_onnx.py:151-161`).  This tool
reads the protobuf directly (no `onnx` package in this image) and writes

    weights/best_model.pth   -- bare 64-key state_dict, the `model_weights.pth`
                               layout of `optimized_train.py:480`, also accepted by
                               `evaluate.py:59-67`

Key recovery: conv / convT / head tensors are named
initializers; the 36 GroupNorm gamma/beta tensors are anonymous `onnx::Mul_N` /
`onnx::Add_N` initializers of shape [C,1,1] whose consuming node output name
(`/enc1/enc1.1/Mul_output_0`) carries the module path.

Usage:  python tools/onnx_weights.py /root/reference/best_model.onnx weights/best_model.pth
"""
import re
import struct
import sys

import numpy as np


def _varint(buf, pos):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _tensor(buf):
    dims, name, raw, floats, dtype = [], None, None, [], None
    for fno, wt, val in _fields(buf):
        if fno == 1:  # dims (maybe packed)
            if wt == 0:
                dims.append(val)
            else:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif fno == 2:
            dtype = val
        elif fno == 4:  # float_data
            if wt == 5:
                floats.append(struct.unpack("<f", val)[0])
            else:
                floats.extend(struct.unpack(f"<{len(val) // 4}f", val))
        elif fno == 8:
            name = bytes(val).decode()
        elif fno == 9:
            raw = bytes(val)
    if dtype != 1:
        return name, None
    if raw is not None:
        arr = np.frombuffer(raw, dtype="<f4").copy()
    else:
        arr = np.asarray(floats, dtype=np.float32)
    return name, arr.reshape(dims)


def _node(buf):
    ins, outs, op = [], [], None
    for fno, _, val in _fields(buf):
        if fno == 1:
            ins.append(bytes(val).decode())
        elif fno == 2:
            outs.append(bytes(val).decode())
        elif fno == 4:
            op = bytes(val).decode()
    return op, ins, outs


def read_onnx_state_dict(path):
    """Return {state_dict key: float32 ndarray} for a LightweightUNet ONNX export."""
    with open(path, "rb") as f:
        model = memoryview(f.read())
    graph = None
    for fno, _, val in _fields(model):
        if fno == 7:
            graph = val
    if graph is None:
        raise ValueError("no GraphProto in file")
    inits, nodes = {}, []
    for fno, _, val in _fields(graph):
        if fno == 5:
            name, arr = _tensor(val)
            if arr is not None:
                inits[name] = arr
        elif fno == 1:
            nodes.append(_node(val))
    state = {}
    for name, arr in inits.items():
        if not name.startswith("onnx::"):
            state[name] = arr
    # GroupNorm affine: Mul -> .weight, Add -> .bias; module path from the node output
    pat = re.compile(r"^/(?P<blk>[\w]+)/(?P<mod>[\w.]+)/(?P<op>Mul|Add)_output_0$")
    for op, ins, outs in nodes:
        if op not in ("Mul", "Add"):
            continue
        anon = [i for i in ins if i.startswith("onnx::") and i in inits]
        if not anon:
            continue
        m = pat.match(outs[0])
        if not m:
            continue
        key = m.group("mod") + (".weight" if op == "Mul" else ".bias")
        state[key] = inits[anon[0]].reshape(-1)
    return state


def main(argv):
    import torch

    src, dst = argv[1], argv[2]
    state = read_onnx_state_dict(src)
    total = sum(v.size for v in state.values())
    print(f"{len(state)} tensors, {total} elements")
    torch.save({k: torch.from_numpy(v) for k, v in state.items()}, dst)
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv)
