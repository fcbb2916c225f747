#!/usr/bin/env python
"""Emit csrc/bg2_tables.h: the shipped 5G base graph as C++ constexpr tables.

The specialised (LDPC_PATH_FAST) kernels are fully unrolled over the base graph at compile
time, so the shift table has to be visible to nvcc as constant expressions.  Input is
codes/bg2_ils0_mod32.triples (3GPP TS 38.212 BG2, set index 0, shifts mod 32; the same data
as the reference's `5G LDPC CODES/NR_2_0_32.txt`, and NR_2_0_4.txt is these values mod 4).
Run by __graft_entry__.build(); the generated header is committed so a plain `nvcc` of the
sources also works.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "codes", "bg2_ils0_mod32.triples")
DST = os.path.join(HERE, "bg2_tables.h")


def read_triples(path):
    rows = cols = None
    cells = []
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        if tok[0] == "rows":
            rows, cols = int(tok[1]), int(tok[3])
            continue
        cells.append(tuple(int(t) for t in tok))
    return rows, cols, cells


def emit(name, rows, cols, cells, Z):
    cells = sorted((i, j, s % Z) for i, j, s in cells)
    coldeg = [0] * cols
    for _, j, _ in cells:
        coldeg[j] += 1
    core_cols = [j for j in range(cols) if coldeg[j] > 1]
    ext_cols = [j for j in range(cols) if coldeg[j] == 1]
    core_of = {j: k for k, j in enumerate(core_cols)}
    ext_of = {j: k for k, j in enumerate(ext_cols)}
    row_ptr = [0]
    col, shift, kind, slot, msgidx = [], [], [], [], []
    ncore_e = 0
    for i in range(rows):
        for (ii, j, s) in cells:
            if ii != i:
                continue
            col.append(j)
            shift.append(s)
            if coldeg[j] > 1:
                kind.append(0)
                slot.append(core_of[j])
                msgidx.append(ncore_e)
                ncore_e += 1
            else:
                kind.append(1)
                slot.append(ext_of[j])
                msgidx.append(-1)
        row_ptr.append(len(col))
    E = len(col)
    maxdc = max(row_ptr[i + 1] - row_ptr[i] for i in range(rows))
    # variable-major views (the edge numbering of create_LLR_mapping, utils/ldpc_utils.py:62-95: edges of variable
    # j*Z+r are consecutive, in ascending check = ascending base row): vm[e] = D_j + k with D_j = edges of the columns
    # before j and k = rank of e's row within column j; cm[e] = the same restricted to core columns (-1 for degree-1)
    col_vm0, acc = [], 0
    for j in range(cols):
        col_vm0.append(acc)
        acc += coldeg[j]
    core_cm0, acc = [], 0
    for j in core_cols:
        core_cm0.append(acc)
        acc += coldeg[j]
    seen = [0] * cols
    vm, cm, vrank = [], [], []
    core_edge = [0] * ncore_e
    ext_edge = [0] * len(ext_cols)
    for e in range(E):
        j = col[e]
        k = seen[j]
        seen[j] += 1
        vm.append(col_vm0[j] + k)
        vrank.append(k)
        if coldeg[j] > 1:
            cm.append(core_cm0[core_of[j]] + k)
            core_edge[core_cm0[core_of[j]] + k] = e
        else:
            cm.append(-1)
            ext_edge[ext_of[j]] = e

    def arr(ctype, nm, vals):
        return f"    static constexpr {ctype} {nm}[{len(vals)}] = {{{', '.join(str(v) for v in vals)}}};"

    out = [f"struct {name} {{",
           f"    static constexpr int kZ = {Z}, kRows = {rows}, kCols = {cols}, kEdges = {E};",
           f"    static constexpr int kCoreCols = {len(core_cols)}, kExtCols = {len(ext_cols)}, kCoreEdges = {ncore_e}, kMaxDc = {maxdc};",
           arr("int", "row_ptr", row_ptr),
           arr("short", "col", col),
           arr("short", "shift", shift),
           arr("signed char", "kind", kind) + "   // 0 = core column (degree > 1), 1 = degree-1 column",
           arr("short", "slot", slot) + "   // index among core columns / among degree-1 columns",
           arr("short", "msg", msgidx) + "   // shared-memory message slot of a core edge, -1 otherwise",
           arr("signed char", "col_kind", [0 if coldeg[j] > 1 else 1 for j in range(cols)]),
           arr("short", "col_slot", [core_of[j] if coldeg[j] > 1 else ext_of[j] for j in range(cols)]),
           arr("short", "core_col", core_cols),
           arr("short", "ext_col", ext_cols),
           arr("short", "col_deg", coldeg),
           arr("short", "col_vm0", col_vm0) + "   // variable-major edge offset of a column (in base edges)",
           arr("short", "vm", vm) + "   // variable-major base-edge index of edge e: col_vm0[col] + rank of its row in the column",
           arr("short", "vrank", vrank),
           arr("short", "cm", cm) + "   // the same among core columns only (-1 for degree-1 columns)",
           arr("short", "core_cm0", core_cm0 + [ncore_e]),
           arr("short", "core_edge", core_edge) + "   // row-major edge index of core-column-major position",
           arr("short", "ext_edge", ext_edge) + "   // row-major edge index of the edge of degree-1 column x",
           "};", ""]
    return "\n".join(out)


def main():
    rows, cols, cells = read_triples(SRC)
    text = ["// GENERATED by csrc/gen_tables.py from codes/bg2_ils0_mod32.triples -- do not edit.",
            "// 5G NR LDPC base graph 2, set index 0 (3GPP TS 38.212 Table 5.3.2-3), as compile-time tables",
            "// for the fully unrolled kernels in decode_fast.cuh.  Edges are row-major, ascending column.",
            "#pragma once", "", "namespace ldpc {", ""]
    text.append(emit("BG2Z32", rows, cols, cells, 32))
    text.append(emit("BG2Z4", rows, cols, cells, 4))
    text.append("}  // namespace ldpc")
    new = "\n".join(text) + "\n"
    old = open(DST).read() if os.path.exists(DST) else None
    if new != old:
        open(DST, "w").write(new)
    return DST


if __name__ == "__main__":
    print(main())
