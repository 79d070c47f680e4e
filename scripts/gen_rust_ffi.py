#!/usr/bin/env python
"""Generates crates-device/rstsr-cuda/src/ffi.rs from include/rstsr_cuda.h: one `extern "C"` declaration per header
prototype, with the header's own argument names.  tests/test_cabi.py checks that the committed ffi.rs equals the
generator's output, so the Rust binding, the ctypes stub and the header cannot drift apart.

    python scripts/gen_rust_ffi.py            # rewrite ffi.rs
    python scripts/gen_rust_ffi.py --print    # print to stdout
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rstsr_cuda.h")
OUT = os.path.join(ROOT, "crates-device", "rstsr-cuda", "src", "ffi.rs")

ENUMS = {"rc_status", "rc_dtype", "rc_order", "rc_iter_order", "rc_binop", "rc_unop", "rc_redop", "rc_uplo", "rc_symm"}
OPAQUE = {"rc_device", "rc_comm"}
SCALARS = {"int": "c_int", "int64_t": "i64", "int32_t": "i32", "uint64_t": "u64", "uint8_t": "u8", "size_t": "usize",
           "double": "f64", "char": "c_char", "void": "c_void"}


def rust_type(ctype: str) -> str:
    t = ctype.strip()
    const = False
    if t.startswith("const "):
        const, t = True, t[6:].strip()
    stars = t.count("*")
    base = t.replace("*", "").strip()
    if base in ENUMS:
        r = "c_int"
    elif base in OPAQUE or base == "rc_layout":
        r = base
    else:
        r = SCALARS[base]
    if stars == 0:
        return r
    out = r
    for i in range(stars):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    return out


def prototypes(text: str):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = text[text.index('extern "C" {') + len('extern "C" {'):]
    for m in re.finditer(r"\n(const char \*|size_t |int )\s*(rc_\w+)\(([^;{]*?)\);", body):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        yield ret, name, args


def convert(ret, name, args):
    params = []
    if args and args != "void":
        for a in args.split(","):
            a = a.strip()
            arr = re.match(r"(.*?)(\w+)\[\w+\]$", a)
            if arr:  # `const uint8_t id[RC_COMM_ID_BYTES]` decays to a pointer
                ctype, pname = arr.group(1).strip() + " *", arr.group(2)
            else:
                m = re.match(r"(.*?)(\w+)$", a)
                ctype, pname = m.group(1).strip(), m.group(2)
            if pname in ("type", "ref", "in", "fn", "move"):
                pname += "_"
            params.append(f"{pname}: {rust_type(ctype)}")
    rret = {"int": "c_int", "size_t": "usize", "const char *": "*const c_char"}[ret]
    line = f"    pub fn {name}({', '.join(params)}) -> {rret};"
    if len(line) > 118:
        inner = ",\n        ".join(params)
        line = f"    pub fn {name}(\n        {inner},\n    ) -> {rret};"
    return line


def generate() -> str:
    text = open(HEADER).read()
    lines = [convert(*p) for p in prototypes(text)]
    head = '''//! `extern "C"` block of librstsr_cuda.so -- GENERATED from include/rstsr_cuda.h by scripts/gen_rust_ffi.py
//! (do not edit; tests/test_cabi.py keeps it in lock-step with the header and with the ctypes stub rstsr_b200/_ffi.py).
//! Enumerations cross the boundary as `c_int`; their Rust-side values live in `crate::codes`.
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_void};

pub const RC_MAX_NDIM: usize = 16;
pub const RC_COMM_ID_BYTES: usize = 128;

/// `Layout<IxD>` as the C side sees it (rstsr-common/src/layout/layoutbase.rs:15-23): element strides, element offset.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rc_layout {
    pub ndim: i32,
    pub shape: [i64; RC_MAX_NDIM],
    pub stride: [i64; RC_MAX_NDIM],
    pub offset: i64,
}

/// Opaque device handle: {ordinal, default order, stream, workspaces}.
#[repr(C)]
pub struct rc_device {
    _private: [u8; 0],
}

/// Opaque communicator: NCCL comm + NVLink peer window.
#[repr(C)]
pub struct rc_comm {
    _private: [u8; 0],
}

#[link(name = "rstsr_cuda")]
extern "C" {
'''
    return head + "\n".join(lines) + "\n}\n"


if __name__ == "__main__":
    src = generate()
    if "--print" in sys.argv:
        sys.stdout.write(src)
    else:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        with open(OUT, "w") as f:
            f.write(src)
        print(OUT, src.count("pub fn"), "functions")
