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
    text.append(emit("BG2Z16", rows, cols, cells, 16))          # the reference's default --lifting_factor (main.py:38)
    text.append(emit("BG2Z8", rows, cols, cells, 8))
    text.append("}  // namespace ldpc")
    new = "\n".join(text) + "\n"
    old = open(DST).read() if os.path.exists(DST) else None
    if new != old:
        open(DST, "w").write(new)
    emit_nq(rows, cols, cells, 32, os.path.join(HERE, "nq_tables.h"))
    return DST




# ---- schedule tables of the rolled QC neural-decoder kernels (csrc/neural_qc_kernel.cuh) ----------------------------------
NQ_MEMBERS = 4                                     # warps that share one codeword (same TMEM lane quarter)
NQ_ROW_CLASSES = [(2, 1), (3, 1), (4, 1), (5, 1), (8, 0), (10, 0)]      # (core edges, degree-1 edges) of a base row
NQ_COL_CLASSES = [6, 8, 10, 13, 16, 23]            # unrolled body sizes for the ordered sums over a base column


def emit_nq(rows, cols, cells, Z, dst):
    cells = sorted((i, j, s % Z) for i, j, s in cells)
    coldeg = [0] * cols
    for _, j, _ in cells:
        coldeg[j] += 1
    core_cols = [j for j in range(cols) if coldeg[j] > 1]
    ext_cols = [j for j in range(cols) if coldeg[j] == 1]
    assert core_cols == list(range(len(core_cols))), "core columns are expected first"
    vm0, acc = [], 0
    for j in range(cols):
        vm0.append(acc)
        acc += coldeg[j]
    seen = [0] * cols
    row_meta, row_ext, row_class = [], [], []
    for i in range(rows):
        meta, ext = [], 255
        for (ii, j, s) in cells:
            if ii != i:
                continue
            k = seen[j]
            seen[j] += 1
            if coldeg[j] > 1:
                meta.append((vm0[j] + k) | (s << 8))
            else:
                assert s == 0 and ext == 255
                ext = ext_cols.index(j)
        row_class.append(NQ_ROW_CLASSES.index((len(meta), 0 if ext == 255 else 1)))
        row_meta.append(meta + [0] * (10 - len(meta)))
        row_ext.append(ext)

    def balance(items, cost, classes_of):
        """LPT: heaviest first to the lightest member; returns per member the items ordered by class."""
        load = [0.0] * NQ_MEMBERS
        got = [[] for _ in range(NQ_MEMBERS)]
        for it in sorted(items, key=lambda x: -cost(x)):
            m = min(range(NQ_MEMBERS), key=lambda q: load[q])
            load[m] += cost(it)
            got[m].append(it)
        return [sorted(g, key=lambda x: (classes_of(x), x)) for g in got], load

    row_cost = lambda i: 31 * NQ_ROW_CLASSES[row_class[i]][0] + 14 * NQ_ROW_CLASSES[row_class[i]][1] + 15
    srows, rload = balance(list(range(rows)), row_cost, lambda i: row_class[i])
    col_class = [min(c for c in range(len(NQ_COL_CLASSES)) if NQ_COL_CLASSES[c] >= coldeg[j]) for j in core_cols]
    col_cost = lambda j: NQ_COL_CLASSES[col_class[j]] * (NQ_COL_CLASSES[col_class[j]] - 1) / 2 + 17 * NQ_COL_CLASSES[col_class[j]] + 10
    scols, cload = balance(core_cols, col_cost, lambda j: col_class[j])
    # degree-1 cells of phase B go to the members with the lightest column load
    sext = [[] for _ in range(NQ_MEMBERS)]
    load = list(cload)
    for x in range(len(ext_cols)):
        m = min(range(NQ_MEMBERS), key=lambda q: load[q])
        load[m] += 12
        sext[m].append(x)
    # backward, variable step: the member's cells flattened (column after column) in chunks of 8, with the member-local slot
    # of their column (its sum is kept in a register); 255 = padding
    scells, sslots = [], []
    for m in range(NQ_MEMBERS):
        cl, sl = [], []
        for slot, j in enumerate(scols[m]):
            for k in range(coldeg[j]):
                cl.append(vm0[j] + k)
                sl.append(slot)
        while len(cl) % 8:
            cl.append(255)
            sl.append(0)
        scells.append(cl)
        sslots.append(sl)
    ncell = max(len(l) for l in scells)
    chunk = []
    cell_col = []
    for m in range(vm0[-1] + coldeg[-1]):
        j = max(jj for jj in range(cols) if vm0[jj] <= m)
        d = coldeg[j]
        inv = 65536 // d + 1
        for off in range(32 * d):
            assert (off * inv) >> 16 == off // d
        chunk.append((vm0[j], d, inv))
        cell_col.append(j)

    def arr(ctype, nm, vals, dims):
        flat = ", ".join(str(v) for v in vals)
        return f"__constant__ {ctype} {nm}{dims} = {{{flat}}};"

    def pad(lists, n):
        return [v for l in lists for v in (l + [0] * (n - len(l)))]

    def prefix(lists, class_of, ncls):
        out = []
        for l in lists:
            cnt = [sum(1 for x in l if class_of(x) == c) for c in range(ncls)]
            p = [0]
            for c in cnt:
                p.append(p[-1] + c)
            out += p
        return out
    nr = max(len(l) for l in srows)
    ncm = max(len(l) for l in scols)
    nxm = max(len(l) for l in sext)
    text = ["// GENERATED by csrc/gen_tables.py (emit_nq) from codes/bg2_ils0_mod32.triples -- do not edit.",
            "// Work schedule of the rolled QC neural-decoder kernels: 4 warps (\"members\") share one codeword; every base row /",
            "// base column is owned by one member; rows are grouped by (core edges, degree-1 edges), columns by body size.",
            f"// member loads (model instructions): rows {[int(x) for x in rload]}, columns + degree-1 cells {[int(x) for x in load]}",
            "#pragma once", "", "namespace ldpc {", "namespace nq {", "",
            f"constexpr int kMembers = {NQ_MEMBERS}, kRowClasses = {len(NQ_ROW_CLASSES)}, kColClasses = {len(NQ_COL_CLASSES)};",
            f"constexpr int kRowsMax = {nr}, kColsMax = {ncm}, kExtMax = {nxm}, kCells = {len(chunk)};",
            arr("unsigned short", "row_meta", pad(row_meta, 10), f"[{rows}][10]") + "   // per core edge of a row: cell | shift << 8",
            arr("unsigned char", "row_ext", row_ext, f"[{rows}]") + "   // degree-1 slot of the row's last edge, 255 = none",
            arr("unsigned char", "sched_rows", pad(srows, nr), f"[{NQ_MEMBERS}][{nr}]"),
            arr("unsigned char", "sched_row_ptr", prefix(srows, lambda i: row_class[i], len(NQ_ROW_CLASSES)), f"[{NQ_MEMBERS}][{len(NQ_ROW_CLASSES) + 1}]"),
            arr("unsigned char", "col_b0", [vm0[j] for j in core_cols], f"[{len(core_cols)}]"),
            arr("unsigned char", "col_d", [coldeg[j] for j in core_cols], f"[{len(core_cols)}]"),
            arr("unsigned char", "sched_cols", pad(scols, ncm), f"[{NQ_MEMBERS}][{ncm}]"),
            arr("unsigned char", "sched_col_ptr", prefix(scols, lambda j: col_class[j], len(NQ_COL_CLASSES)), f"[{NQ_MEMBERS}][{len(NQ_COL_CLASSES) + 1}]"),
            f"constexpr int kCellsMax = {ncell};",
            arr("unsigned char", "sched_cells", [v for l in scells for v in (l + [255] * (ncell - len(l)))], f"[{NQ_MEMBERS}][{ncell}]"),
            arr("unsigned char", "sched_cell_slot", [v for l in sslots for v in (l + [0] * (ncell - len(l)))], f"[{NQ_MEMBERS}][{ncell}]"),
            arr("unsigned char", "sched_cell_cnt", [len(l) for l in scells], f"[{NQ_MEMBERS}]"),
            arr("unsigned char", "sched_ext", pad(sext, nxm), f"[{NQ_MEMBERS}][{nxm}]"),
            arr("unsigned char", "sched_ext_cnt", [len(l) for l in sext], f"[{NQ_MEMBERS}]"),
            arr("unsigned int", "chunk_meta", [D | (d << 8) | (inv << 13) for D, d, inv in chunk], f"[{len(chunk)}]") +
            "   // per 32-edge chunk of a row: first cell of its column | degree << 8 | (65536 / degree + 1) << 13",
            arr("unsigned char", "cell_col", cell_col, f"[{len(cell_col)}]") + "   // base column of a cell (per-variable I/O)",
            "", "}  // namespace nq", "}  // namespace ldpc", ""]
    new = "\n".join(text)
    old = open(dst).read() if os.path.exists(dst) else None
    if new != old:
        open(dst, "w").write(new)
    return rload, load


if __name__ == "__main__":
    print(main())
