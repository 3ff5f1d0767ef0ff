/*
 * pyrows.c -- CPython helper for the host side of the drop-in: turns the numeric rows that come back
 * from libpavgpu.so into the Python objects pandas needs (object-dtype columns of str / int), the one
 * part of make_insdel_snv_calls that cannot leave the interpreter. Replaces per-row f-strings
 * (reference: pavlib/cigarcall.py:112,121,185,198,254,267 build the same strings one pd.Series at a time).
 *
 *   format(n, parts)  -> object ndarray of str   parts = sequence of ('s', str) | ('i', int64 buffer) |
 *                                            ('l', list[str], int64 index buffer) | ('c', uint8 buffer)
 *   ints(int64 buffer) -> object ndarray of int
 *   slices(data uint8 buffer list, which int64 buffer, start int64 buffer, length int64 buffer,
 *          rc uint8 buffer, comp bytes[256]) -> object ndarray of str   (reverse-complemented through comp when rc[i])
 *
 * Host-only; no CUDA here. Built in-tree by pav_b200/build.py with gcc.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <stdint.h>
#include <string.h>

/* Results are 1-D numpy object arrays filled in place (no intermediate list, no type inference). */
static PyObject *new_obj_array(Py_ssize_t n, PyObject ***data)
{
    npy_intp dims[1] = {(npy_intp)n};
    PyObject *arr = PyArray_SimpleNew(1, dims, NPY_OBJECT);   /* object arrays come back NULL-filled */
    if (!arr) return NULL;
    *data = (PyObject **)PyArray_DATA((PyArrayObject *)arr);
    return arr;
}

enum { P_LIT, P_INT, P_LUT, P_CHR };

typedef struct {
    int kind;
    const char *lit; Py_ssize_t lit_len;
    Py_buffer buf; int has_buf;
    PyObject *lut;                 /* borrowed list */
    const char **lut_s; Py_ssize_t *lut_n; Py_ssize_t lut_size;
} part_t;

static void parts_free(part_t *p, Py_ssize_t n)
{
    for (Py_ssize_t i = 0; i < n; i++) {
        if (p[i].has_buf) PyBuffer_Release(&p[i].buf);
        PyMem_Free(p[i].lut_s); PyMem_Free(p[i].lut_n);
    }
    PyMem_Free(p);
}

static inline char *put_i64(char *dst, int64_t v)
{
    char tmp[24]; int k = 0;
    uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    if (v < 0) *dst++ = '-';
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (k) *dst++ = tmp[--k];
    return dst;
}

static PyObject *py_format(PyObject *self, PyObject *args)
{
    Py_ssize_t n; PyObject *parts_obj;
    if (!PyArg_ParseTuple(args, "nO", &n, &parts_obj)) return NULL;
    PyObject *seq = PySequence_Fast(parts_obj, "parts must be a sequence");
    if (!seq) return NULL;
    Py_ssize_t np_ = PySequence_Fast_GET_SIZE(seq);
    part_t *parts = (part_t *)PyMem_Calloc((size_t)np_ + 1, sizeof(part_t));
    if (!parts) { Py_DECREF(seq); return PyErr_NoMemory(); }
    Py_ssize_t max_len = 1;
    for (Py_ssize_t i = 0; i < np_; i++) {
        PyObject *it = PySequence_Fast_GET_ITEM(seq, i);
        if (!PyTuple_Check(it) || PyTuple_GET_SIZE(it) < 2) { PyErr_SetString(PyExc_TypeError, "part must be a tuple"); goto fail; }
        const char *kind = PyUnicode_AsUTF8(PyTuple_GET_ITEM(it, 0));
        if (!kind) goto fail;
        part_t *p = &parts[i];
        if (kind[0] == 's') {
            p->kind = P_LIT;
            p->lit = PyUnicode_AsUTF8AndSize(PyTuple_GET_ITEM(it, 1), &p->lit_len);
            if (!p->lit) goto fail;
            max_len += p->lit_len;
        } else if (kind[0] == 'i' || kind[0] == 'c') {
            p->kind = kind[0] == 'i' ? P_INT : P_CHR;
            if (PyObject_GetBuffer(PyTuple_GET_ITEM(it, 1), &p->buf, PyBUF_SIMPLE) < 0) goto fail;
            p->has_buf = 1;
            Py_ssize_t need = n * (p->kind == P_INT ? 8 : 1);
            if (p->buf.len < need) { PyErr_SetString(PyExc_ValueError, "buffer shorter than n rows"); goto fail; }
            max_len += p->kind == P_INT ? 21 : 1;
        } else if (kind[0] == 'l') {
            if (PyTuple_GET_SIZE(it) < 3 || !PyList_Check(PyTuple_GET_ITEM(it, 1))) { PyErr_SetString(PyExc_TypeError, "('l', list, index buffer)"); goto fail; }
            p->kind = P_LUT; p->lut = PyTuple_GET_ITEM(it, 1);
            p->lut_size = PyList_GET_SIZE(p->lut);
            p->lut_s = (const char **)PyMem_Calloc((size_t)p->lut_size + 1, sizeof(char *));
            p->lut_n = (Py_ssize_t *)PyMem_Calloc((size_t)p->lut_size + 1, sizeof(Py_ssize_t));
            if (!p->lut_s || !p->lut_n) { PyErr_NoMemory(); goto fail; }
            Py_ssize_t mx = 0;
            for (Py_ssize_t k = 0; k < p->lut_size; k++) {
                p->lut_s[k] = PyUnicode_AsUTF8AndSize(PyList_GET_ITEM(p->lut, k), &p->lut_n[k]);
                if (!p->lut_s[k]) goto fail;
                if (p->lut_n[k] > mx) mx = p->lut_n[k];
            }
            if (PyObject_GetBuffer(PyTuple_GET_ITEM(it, 2), &p->buf, PyBUF_SIMPLE) < 0) goto fail;
            p->has_buf = 1;
            if (p->buf.len < n * 8) { PyErr_SetString(PyExc_ValueError, "index buffer shorter than n rows"); goto fail; }
            max_len += mx;
        } else { PyErr_SetString(PyExc_ValueError, "unknown part kind"); goto fail; }
    }
    {
        char *scratch = (char *)PyMem_Malloc((size_t)max_len + 8);
        PyObject **slots = NULL;
        PyObject *out = new_obj_array(n, &slots);
        if (!scratch || !out) { PyMem_Free(scratch); Py_XDECREF(out); PyErr_NoMemory(); goto fail; }
        for (Py_ssize_t r = 0; r < n; r++) {
            char *d = scratch;
            for (Py_ssize_t i = 0; i < np_; i++) {
                part_t *p = &parts[i];
                switch (p->kind) {
                case P_LIT: memcpy(d, p->lit, (size_t)p->lit_len); d += p->lit_len; break;
                case P_INT: d = put_i64(d, ((const int64_t *)p->buf.buf)[r]); break;
                case P_CHR: *d++ = (char)((const uint8_t *)p->buf.buf)[r]; break;
                case P_LUT: {
                    int64_t k = ((const int64_t *)p->buf.buf)[r];
                    if (k < 0 || k >= p->lut_size) { PyMem_Free(scratch); Py_DECREF(out); PyErr_SetString(PyExc_IndexError, "lookup index out of range"); goto fail; }
                    memcpy(d, p->lut_s[k], (size_t)p->lut_n[k]); d += p->lut_n[k];
                    break; }
                }
            }
            /* all pieces are ASCII in practice (names, digits, base letters): build the compact str directly */
            Py_ssize_t len = d - scratch;
            int ascii = 1;
            for (Py_ssize_t k = 0; k < len; k++) if ((unsigned char)scratch[k] & 0x80) { ascii = 0; break; }
            PyObject *s;
            if (ascii) {
                s = PyUnicode_New(len, 127);
                if (s) memcpy(PyUnicode_1BYTE_DATA(s), scratch, (size_t)len);
            } else {
                s = PyUnicode_DecodeUTF8(scratch, len, NULL);
            }
            if (!s) { PyMem_Free(scratch); Py_DECREF(out); goto fail; }
            slots[r] = s;
        }
        PyMem_Free(scratch);
        parts_free(parts, np_);
        Py_DECREF(seq);
        return out;
    }
fail:
    parts_free(parts, np_);
    Py_DECREF(seq);
    return NULL;
}

static PyObject *py_ints(PyObject *self, PyObject *arg)
{
    Py_buffer b;
    if (PyObject_GetBuffer(arg, &b, PyBUF_SIMPLE) < 0) return NULL;
    Py_ssize_t n = b.len / 8;
    PyObject **slots = NULL;
    PyObject *out = new_obj_array(n, &slots);
    if (out) {
        const int64_t *v = (const int64_t *)b.buf;
        for (Py_ssize_t i = 0; i < n; i++) {
            PyObject *o = PyLong_FromLongLong(v[i]);
            if (!o) { Py_DECREF(out); out = NULL; break; }
            slots[i] = o;
        }
    }
    PyBuffer_Release(&b);
    return out;
}

static PyObject *py_slices(PyObject *self, PyObject *args)
{
    PyObject *data_list; Py_buffer which, start, length, rc, comp;
    if (!PyArg_ParseTuple(args, "O!y*y*y*y*y*", &PyList_Type, &data_list, &which, &start, &length, &rc, &comp)) return NULL;
    PyObject *out = NULL;
    PyObject **slots = NULL;
    Py_ssize_t n = which.len / 8, nd = PyList_GET_SIZE(data_list);
    Py_buffer *bufs = (Py_buffer *)PyMem_Calloc((size_t)nd + 1, sizeof(Py_buffer));
    Py_ssize_t got = 0;
    char *scratch = NULL; Py_ssize_t cap = 0;
    if (!bufs) { PyErr_NoMemory(); goto done; }
    if (start.len < n * 8 || length.len < n * 8 || rc.len < n || comp.len < 256) { PyErr_SetString(PyExc_ValueError, "slices: bad buffer sizes"); goto done; }
    for (; got < nd; got++)
        if (PyObject_GetBuffer(PyList_GET_ITEM(data_list, got), &bufs[got], PyBUF_SIMPLE) < 0) goto done;
    out = new_obj_array(n, &slots);
    if (!out) goto done;
    for (Py_ssize_t i = 0; i < n; i++) {
        int64_t w = ((const int64_t *)which.buf)[i], s = ((const int64_t *)start.buf)[i], l = ((const int64_t *)length.buf)[i];
        if (w < 0 || w >= nd || s < 0 || l < 0 || s + l > bufs[w].len) { Py_CLEAR(out); PyErr_SetString(PyExc_IndexError, "slices: range outside sequence"); goto done; }
        const uint8_t *src = (const uint8_t *)bufs[w].buf + s;
        PyObject *o;
        if (((const uint8_t *)rc.buf)[i]) {
            if (l > cap) { PyMem_Free(scratch); cap = l * 2 + 64; scratch = (char *)PyMem_Malloc((size_t)cap); if (!scratch) { Py_CLEAR(out); PyErr_NoMemory(); goto done; } }
            const uint8_t *ct = (const uint8_t *)comp.buf;
            for (int64_t k = 0; k < l; k++) scratch[k] = (char)ct[src[l - 1 - k]];
            o = PyUnicode_DecodeLatin1(scratch, l, NULL);
        } else {
            o = PyUnicode_DecodeLatin1((const char *)src, l, NULL);
        }
        if (!o) { Py_CLEAR(out); goto done; }
        slots[i] = o;
    }
done:
    PyMem_Free(scratch);
    for (Py_ssize_t k = 0; k < got; k++) PyBuffer_Release(&bufs[k]);
    PyMem_Free(bufs);
    PyBuffer_Release(&which); PyBuffer_Release(&start); PyBuffer_Release(&length); PyBuffer_Release(&rc); PyBuffer_Release(&comp);
    return out;
}

/* ------------------------------------------------------------------------------------------------
 * Whole-frame builders: every column of df_snv / df_insdel in one pass over the device rows, already in
 * the final (sorted) row order. Strings are written straight into exactly-sized compact ASCII str objects
 * (length from digit counts; no scratch copy), shared values (chromosome, strand, HAP, SVTYPE, small
 * "l,r" homology strings, single bases) are one object referenced many times.
 * ------------------------------------------------------------------------------------------------ */
static const char DIG2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
    "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
    "8081828384858687888990919293949596979899";

static inline int ndig_u64(uint64_t v)
{
    int n = 1;
    for (;;) {
        if (v < 10) return n;
        if (v < 100) return n + 1;
        if (v < 1000) return n + 2;
        if (v < 10000) return n + 3;
        v /= 10000u; n += 4;
    }
}

/* digits of v, most significant first, ending just before `end`; returns the first written byte */
static inline char *put_u64_back(char *end, uint64_t v)
{
    while (v >= 100) { unsigned r = (unsigned)(v % 100); v /= 100; end -= 2; memcpy(end, DIG2 + 2 * r, 2); }
    if (v >= 10) { end -= 2; memcpy(end, DIG2 + 2 * v, 2); }
    else *--end = (char)('0' + v);
    return end;
}

static inline int ndig_i64(int64_t v) { return v < 0 ? 1 + ndig_u64((uint64_t)(-(v + 1)) + 1u) : ndig_u64((uint64_t)v); }

static inline char *put_i64_fwd(char *dst, int64_t v, int nd)   /* nd = ndig_i64(v) */
{
    uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    if (v < 0) *dst = '-';
    put_u64_back(dst + nd, u);
    return dst + nd;
}

typedef struct {   /* ASCII strings of a list, or failure */
    Py_ssize_t n;
    const char **s;
    Py_ssize_t *len;
} strtab_t;

static int strtab_load(strtab_t *t, PyObject *list, const char *what)
{
    t->n = 0; t->s = NULL; t->len = NULL;
    if (!PyList_Check(list)) { PyErr_Format(PyExc_TypeError, "%s must be a list of str", what); return -1; }
    Py_ssize_t n = PyList_GET_SIZE(list);
    t->s = (const char **)PyMem_Calloc((size_t)n + 1, sizeof(char *));
    t->len = (Py_ssize_t *)PyMem_Calloc((size_t)n + 1, sizeof(Py_ssize_t));
    if (!t->s || !t->len) { PyErr_NoMemory(); return -1; }
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *o = PyList_GET_ITEM(list, i);
        if (!PyUnicode_Check(o) || !PyUnicode_IS_ASCII(o)) { PyErr_Format(PyExc_ValueError, "%s: entry %zd is not an ASCII str", what, i); return -1; }
        t->s[i] = (const char *)PyUnicode_1BYTE_DATA(o);
        t->len[i] = PyUnicode_GET_LENGTH(o);
    }
    t->n = n;
    return 0;
}

static void strtab_free(strtab_t *t) { PyMem_Free(t->s); PyMem_Free(t->len); }

static int list_of(PyObject *o, Py_ssize_t n, const char *what)
{
    if (!PyList_Check(o) || PyList_GET_SIZE(o) < n) { PyErr_Format(PyExc_TypeError, "%s must be a list with one entry per record", what); return -1; }
    return 0;
}

static inline void fill_const(PyObject **slots, Py_ssize_t n, PyObject *v)
{
    for (Py_ssize_t i = 0; i < n; i++) { Py_INCREF(v); slots[i] = v; }
}

/* REF / ALT of consecutive rows sit ~100 bases apart in two multi-hundred-MB arrays: one cache miss each per row unless the
 * lines are requested a few rows ahead. */
#define PREFETCH_AHEAD 16
#define SNV_NCOL 14
#define INDEL_NCOL 16

typedef struct { int32_t pos_ref, qry_pos, rec, op_idx; } snv_row_t;
typedef struct { int32_t rec, op_idx, svtype, svlen, pos, end, qry_pos, qry_end, left_shift, hom_ref_l, hom_ref_r, hom_tig_l, hom_tig_r, seq_start, pad[2]; } indel_row_t;

static inline void prefetch_snv(const snv_row_t *R, const int64_t *ord, Py_ssize_t k, Py_ssize_t n, Py_ssize_t n_rows, Py_ssize_t n_rec,
                                const int32_t *rid, const int32_t *qid, const Py_buffer *bufs, Py_ssize_t n_ref, Py_ssize_t nd_seq)
{
    if (k + PREFETCH_AHEAD >= n) return;
    const int64_t i = ord[k + PREFETCH_AHEAD];
    if (i < 0 || i >= n_rows) return;
    const snv_row_t *r = &R[i];
    __builtin_prefetch(r + 8);
    if (r->rec < 0 || r->rec >= n_rec) return;
    const Py_ssize_t wr = rid[r->rec], wq = n_ref + (Py_ssize_t)qid[r->rec];
    if (wr >= 0 && wr < n_ref && r->pos_ref >= 0 && r->pos_ref < bufs[wr].len) __builtin_prefetch((const char *)bufs[wr].buf + r->pos_ref);
    if (wq >= n_ref && wq < nd_seq && r->qry_pos >= 0 && r->qry_pos < bufs[wq].len) __builtin_prefetch((const char *)bufs[wq].buf + r->qry_pos);
}

static PyObject *one_char_table[256];

static PyObject *one_char(unsigned c)
{
    if (!one_char_table[c]) {
        char ch = (char)c;
        one_char_table[c] = PyUnicode_DecodeLatin1(&ch, 1, NULL);
    }
    return one_char_table[c];
}

/* snv_frame(rows, order, ids|None, chrom_objs, chrom_strs, qry_strs, strand_objs, align_index_objs, ref_id, qry_id, rev,
 *           seqs, n_ref, comp, consts)
 *   rows: pavgpu_snv_row buffer (emission order); order: int64 permutation (final row k shows emission row order[k]);
 *   ids: object ndarray of ready IDs in emission order, or None to format "{chrom}-{pos+1}-SNV-{REF}{ALT}" (upper-cased bases)
 *   here; ref_id / qry_id: int32 per record; rev: uint8 per record; seqs: list of uint8 buffers = reference sequences then
 *   contigs (forward strand): REF = reference[pos], ALT = contig[qry_pos], complemented through comp for minus-strand records
 *   (cigarcall.py:69-70,105-106); consts = (svtype 'SNV', svlen 1, hap, ci 0, call_source)
 * -> tuple of 14 object ndarrays in column order of the reference (pavlib/cigarcall.py:125-134). */
static PyObject *py_snv_frame(PyObject *self, PyObject *args)
{
    Py_buffer rows, order, ref_id, qry_id, rev, comp;
    PyObject *ids, *chrom_objs, *chrom_strs, *qry_strs, *strand_objs, *ai_objs, *seqs, *consts;
    Py_ssize_t n_ref;
    if (!PyArg_ParseTuple(args, "y*y*OOOOOOy*y*y*O!ny*O", &rows, &order, &ids, &chrom_objs, &chrom_strs, &qry_strs, &strand_objs, &ai_objs,
                          &ref_id, &qry_id, &rev, &PyList_Type, &seqs, &n_ref, &comp, &consts))
        return NULL;
    PyObject *result = NULL, *cols[SNV_NCOL] = {0};
    PyObject **slot[SNV_NCOL];
    strtab_t tc = {0}, tq = {0};
    PyObject **id_src = NULL;
    Py_ssize_t n = order.len / 8, n_rows = rows.len / (Py_ssize_t)sizeof(snv_row_t), nd_seq = PyList_GET_SIZE(seqs), got = 0;
    Py_buffer *bufs = (Py_buffer *)PyMem_Calloc((size_t)nd_seq + 1, sizeof(Py_buffer));
    if (!bufs) { PyErr_NoMemory(); goto done; }
    if (n > n_rows || comp.len < 256) { PyErr_SetString(PyExc_ValueError, "snv_frame: buffer sizes disagree"); goto done; }
    if (strtab_load(&tc, chrom_strs, "chrom_strs") < 0 || strtab_load(&tq, qry_strs, "qry_strs") < 0) goto done;
    Py_ssize_t n_rec = tc.n;
    if (tq.n < n_rec || ref_id.len < n_rec * 4 || qry_id.len < n_rec * 4 || rev.len < n_rec || list_of(chrom_objs, n_rec, "chrom_objs") < 0 ||
        list_of(strand_objs, n_rec, "strand_objs") < 0 || list_of(ai_objs, n_rec, "align_index_objs") < 0) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "snv_frame: per-record arrays disagree");
        goto done;
    }
    if (!PyTuple_Check(consts) || PyTuple_GET_SIZE(consts) != 5) { PyErr_SetString(PyExc_TypeError, "snv_frame: consts must be a 5-tuple"); goto done; }
    if (ids != Py_None) {
        if (!PyArray_Check(ids) || PyArray_TYPE((PyArrayObject *)ids) != NPY_OBJECT || PyArray_NDIM((PyArrayObject *)ids) != 1 ||
            PyArray_DIM((PyArrayObject *)ids, 0) < n_rows || !PyArray_IS_C_CONTIGUOUS((PyArrayObject *)ids)) {
            PyErr_SetString(PyExc_TypeError, "snv_frame: ids must be a contiguous 1-D object array with one entry per row"); goto done;
        }
        id_src = (PyObject **)PyArray_DATA((PyArrayObject *)ids);
    }
    for (; got < nd_seq; got++)
        if (PyObject_GetBuffer(PyList_GET_ITEM(seqs, got), &bufs[got], PyBUF_SIMPLE) < 0) goto done;
    for (int c = 0; c < SNV_NCOL; c++) { cols[c] = new_obj_array(n, &slot[c]); if (!cols[c]) goto done; }
    {
        const snv_row_t *R = (const snv_row_t *)rows.buf;
        const int64_t *ord = (const int64_t *)order.buf;
        const int32_t *rid = (const int32_t *)ref_id.buf, *qid = (const int32_t *)qry_id.buf;
        const uint8_t *rv = (const uint8_t *)rev.buf, *ct = (const uint8_t *)comp.buf;
        PyObject *c_svtype = PyTuple_GET_ITEM(consts, 0), *c_svlen = PyTuple_GET_ITEM(consts, 1), *c_hap = PyTuple_GET_ITEM(consts, 2),
                 *c_ci = PyTuple_GET_ITEM(consts, 3), *c_src = PyTuple_GET_ITEM(consts, 4);
        fill_const(slot[4], n, c_svtype); fill_const(slot[5], n, c_svlen); fill_const(slot[8], n, c_hap);
        fill_const(slot[11], n, c_ci); fill_const(slot[13], n, c_src);
        for (Py_ssize_t k = 0; k < n; k++) {
            int64_t i = ord[k];
            prefetch_snv(R, ord, k, n, n_rows, n_rec, rid, qid, bufs, n_ref, nd_seq);
            if (i < 0 || i >= n_rows) { PyErr_SetString(PyExc_IndexError, "snv_frame: order entry out of range"); goto done; }
            const snv_row_t r = R[i];
            if (r.rec < 0 || r.rec >= n_rec) { PyErr_SetString(PyExc_IndexError, "snv_frame: record index out of range"); goto done; }
            const Py_ssize_t wr = rid[r.rec], wq = n_ref + (Py_ssize_t)qid[r.rec];
            if (wr < 0 || wr >= n_ref || wq < n_ref || wq >= nd_seq || r.pos_ref < 0 || r.pos_ref >= bufs[wr].len || r.qry_pos < 0 || r.qry_pos >= bufs[wq].len) {
                PyErr_SetString(PyExc_IndexError, "snv_frame: position outside its sequence"); goto done;
            }
            const unsigned rb = ((const uint8_t *)bufs[wr].buf)[r.pos_ref];
            unsigned ab = ((const uint8_t *)bufs[wq].buf)[r.qry_pos];
            if (rv[r.rec]) ab = ct[ab];
            PyObject *o;
            o = PyList_GET_ITEM(chrom_objs, r.rec); Py_INCREF(o); slot[0][k] = o;
            if (!(slot[1][k] = PyLong_FromLong(r.pos_ref))) goto done;
            if (!(slot[2][k] = PyLong_FromLong((long)r.pos_ref + 1))) goto done;
            if (id_src) { o = id_src[i]; Py_INCREF(o); slot[3][k] = o; }
            else {   /* {chrom}-{pos+1}-SNV-{REF}{ALT} */
                if ((rb | ab) & 0x80) { PyErr_SetString(PyExc_ValueError, "snv_frame: non-ASCII base"); goto done; }
                int64_t p1 = (int64_t)r.pos_ref + 1;
                int nd = ndig_i64(p1);
                Py_ssize_t cl = tc.len[r.rec], len = cl + 1 + nd + 5 + 2;
                if (!(o = PyUnicode_New(len, 127))) goto done;
                char *d = (char *)PyUnicode_1BYTE_DATA(o);
                memcpy(d, tc.s[r.rec], (size_t)cl); d += cl;
                *d++ = '-'; d = put_i64_fwd(d, p1, nd);
                memcpy(d, "-SNV-", 5); d += 5;
                *d++ = (char)((rb >= 'a' && rb <= 'z') ? rb - 32 : rb);
                *d++ = (char)((ab >= 'a' && ab <= 'z') ? ab - 32 : ab);
                slot[3][k] = o;
            }
            if (!(o = one_char(rb))) goto done;
            Py_INCREF(o); slot[6][k] = o;
            if (!(o = one_char(ab))) goto done;
            Py_INCREF(o); slot[7][k] = o;
            {   /* {qry}:{qp+1}-{qp+1} */
                int64_t q1 = (int64_t)r.qry_pos + 1;
                int nd = ndig_i64(q1);
                Py_ssize_t ql = tq.len[r.rec], len = ql + 1 + nd + 1 + nd;
                if (!(o = PyUnicode_New(len, 127))) goto done;
                char *d = (char *)PyUnicode_1BYTE_DATA(o);
                memcpy(d, tq.s[r.rec], (size_t)ql); d += ql;
                *d++ = ':'; d = put_i64_fwd(d, q1, nd);
                *d++ = '-'; d = put_i64_fwd(d, q1, nd);
                slot[9][k] = o;
            }
            o = PyList_GET_ITEM(strand_objs, r.rec); Py_INCREF(o); slot[10][k] = o;
            o = PyList_GET_ITEM(ai_objs, r.rec); Py_INCREF(o); slot[12][k] = o;
        }
    }
    result = PyTuple_New(SNV_NCOL);
    if (result) for (int c = 0; c < SNV_NCOL; c++) { PyTuple_SET_ITEM(result, c, cols[c]); cols[c] = NULL; }
done:
    for (int c = 0; c < SNV_NCOL; c++) Py_XDECREF(cols[c]);
    if (bufs) { for (Py_ssize_t j = 0; j < got; j++) PyBuffer_Release(&bufs[j]); PyMem_Free(bufs); }
    strtab_free(&tc); strtab_free(&tq);
    PyBuffer_Release(&rows); PyBuffer_Release(&order); PyBuffer_Release(&ref_id); PyBuffer_Release(&qry_id); PyBuffer_Release(&rev); PyBuffer_Release(&comp);
    return result;
}

#define HOM_CACHE 64
static PyObject *hom_cache[HOM_CACHE][HOM_CACHE];

static PyObject *hom_str(int32_t l, int32_t r)   /* new reference to "l,r" */
{
    int cached = l >= 0 && r >= 0 && l < HOM_CACHE && r < HOM_CACHE;
    if (cached && hom_cache[l][r]) { Py_INCREF(hom_cache[l][r]); return hom_cache[l][r]; }
    int nl = ndig_i64(l), nr = ndig_i64(r);
    PyObject *o = PyUnicode_New(nl + 1 + nr, 127);
    if (!o) return NULL;
    char *d = (char *)PyUnicode_1BYTE_DATA(o);
    d = put_i64_fwd(d, l, nl); *d++ = ','; put_i64_fwd(d, r, nr);
    if (cached) { hom_cache[l][r] = o; Py_INCREF(o); }
    return o;
}

/* indel_frame(rows, order, ids|None, chrom_objs, chrom_strs, qry_strs, strand_objs, align_index_objs, ref_id, qry_id, rev,
 *             seqs, n_ref, comp, consts)
 *   rows: pavgpu_indel_row buffer (emission order); ref_id / qry_id: int32 per record; rev: uint8 per record;
 *   seqs: list of uint8 buffers = reference sequences then contigs (forward strand); comp: 256-byte complement table;
 *   consts = ('INS', 'DEL', hap, ci 0, call_source)
 * -> tuple of 16 object ndarrays in column order of the reference (pavlib/cigarcall.py:199-209). */
static PyObject *py_indel_frame(PyObject *self, PyObject *args)
{
    Py_buffer rows, order, ref_id, qry_id, rev, comp;
    PyObject *ids, *chrom_objs, *chrom_strs, *qry_strs, *strand_objs, *ai_objs, *seqs, *consts;
    Py_ssize_t n_ref;
    if (!PyArg_ParseTuple(args, "y*y*OOOOOOy*y*y*O!ny*O", &rows, &order, &ids, &chrom_objs, &chrom_strs, &qry_strs, &strand_objs, &ai_objs,
                          &ref_id, &qry_id, &rev, &PyList_Type, &seqs, &n_ref, &comp, &consts))
        return NULL;
    PyObject *result = NULL, *cols[INDEL_NCOL] = {0};
    PyObject **slot[INDEL_NCOL];
    strtab_t tc = {0}, tq = {0};
    PyObject **id_src = NULL;
    Py_ssize_t n = order.len / 8, n_rows = rows.len / (Py_ssize_t)sizeof(indel_row_t), nd_seq = PyList_GET_SIZE(seqs), got = 0;
    Py_buffer *bufs = (Py_buffer *)PyMem_Calloc((size_t)nd_seq + 1, sizeof(Py_buffer));
    char *scratch = NULL; Py_ssize_t cap = 0;
    if (!bufs) { PyErr_NoMemory(); goto done; }
    if (n > n_rows || comp.len < 256) { PyErr_SetString(PyExc_ValueError, "indel_frame: buffer sizes disagree"); goto done; }
    if (strtab_load(&tc, chrom_strs, "chrom_strs") < 0 || strtab_load(&tq, qry_strs, "qry_strs") < 0) goto done;
    Py_ssize_t n_rec = tc.n;
    if (tq.n < n_rec || ref_id.len < n_rec * 4 || qry_id.len < n_rec * 4 || rev.len < n_rec || list_of(chrom_objs, n_rec, "chrom_objs") < 0 ||
        list_of(strand_objs, n_rec, "strand_objs") < 0 || list_of(ai_objs, n_rec, "align_index_objs") < 0) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "indel_frame: per-record arrays disagree");
        goto done;
    }
    if (!PyTuple_Check(consts) || PyTuple_GET_SIZE(consts) != 5) { PyErr_SetString(PyExc_TypeError, "indel_frame: consts must be a 5-tuple"); goto done; }
    if (ids != Py_None) {
        if (!PyArray_Check(ids) || PyArray_TYPE((PyArrayObject *)ids) != NPY_OBJECT || PyArray_NDIM((PyArrayObject *)ids) != 1 ||
            PyArray_DIM((PyArrayObject *)ids, 0) < n_rows || !PyArray_IS_C_CONTIGUOUS((PyArrayObject *)ids)) {
            PyErr_SetString(PyExc_TypeError, "indel_frame: ids must be a contiguous 1-D object array with one entry per row"); goto done;
        }
        id_src = (PyObject **)PyArray_DATA((PyArrayObject *)ids);
    }
    for (; got < nd_seq; got++)
        if (PyObject_GetBuffer(PyList_GET_ITEM(seqs, got), &bufs[got], PyBUF_SIMPLE) < 0) goto done;
    for (int c = 0; c < INDEL_NCOL; c++) { cols[c] = new_obj_array(n, &slot[c]); if (!cols[c]) goto done; }
    {
        const indel_row_t *R = (const indel_row_t *)rows.buf;
        const int64_t *ord = (const int64_t *)order.buf;
        const int32_t *rid = (const int32_t *)ref_id.buf, *qid = (const int32_t *)qry_id.buf;
        const uint8_t *rv = (const uint8_t *)rev.buf, *ct = (const uint8_t *)comp.buf;
        PyObject *c_ins = PyTuple_GET_ITEM(consts, 0), *c_del = PyTuple_GET_ITEM(consts, 1), *c_hap = PyTuple_GET_ITEM(consts, 2),
                 *c_ci = PyTuple_GET_ITEM(consts, 3), *c_src = PyTuple_GET_ITEM(consts, 4);
        fill_const(slot[6], n, c_hap); fill_const(slot[9], n, c_ci); fill_const(slot[14], n, c_src);
        for (Py_ssize_t k = 0; k < n; k++) {
            int64_t i = ord[k];
            if (i < 0 || i >= n_rows) { PyErr_SetString(PyExc_IndexError, "indel_frame: order entry out of range"); goto done; }
            const indel_row_t r = R[i];
            if (r.rec < 0 || r.rec >= n_rec) { PyErr_SetString(PyExc_IndexError, "indel_frame: record index out of range"); goto done; }
            const int is_del = r.svtype == 1;
            PyObject *o;
            o = PyList_GET_ITEM(chrom_objs, r.rec); Py_INCREF(o); slot[0][k] = o;
            if (!(slot[1][k] = PyLong_FromLong(r.pos))) goto done;
            if (!(slot[2][k] = PyLong_FromLong(r.end))) goto done;
            if (id_src) { o = id_src[i]; Py_INCREF(o); slot[3][k] = o; }
            else {   /* {chrom}-{pos+1}-{INS|DEL}-{svlen} */
                int64_t p1 = (int64_t)r.pos + 1;
                int nd = ndig_i64(p1), nl = ndig_i64(r.svlen);
                Py_ssize_t cl = tc.len[r.rec], len = cl + 1 + nd + 5 + nl;
                if (!(o = PyUnicode_New(len, 127))) goto done;
                char *d = (char *)PyUnicode_1BYTE_DATA(o);
                memcpy(d, tc.s[r.rec], (size_t)cl); d += cl;
                *d++ = '-'; d = put_i64_fwd(d, p1, nd);
                memcpy(d, is_del ? "-DEL-" : "-INS-", 5); d += 5;
                put_i64_fwd(d, r.svlen, nl);
                slot[3][k] = o;
            }
            o = is_del ? c_del : c_ins; Py_INCREF(o); slot[4][k] = o;
            if (!(slot[5][k] = PyLong_FromLong(r.svlen))) goto done;
            {   /* {qry}:{qp+1}-{qe}  (DEL: qe = qp+1) */
                int64_t q1 = (int64_t)r.qry_pos + 1, q2 = is_del ? q1 : (int64_t)r.qry_end;
                int n1 = ndig_i64(q1), n2 = ndig_i64(q2);
                Py_ssize_t ql = tq.len[r.rec], len = ql + 1 + n1 + 1 + n2;
                if (!(o = PyUnicode_New(len, 127))) goto done;
                char *d = (char *)PyUnicode_1BYTE_DATA(o);
                memcpy(d, tq.s[r.rec], (size_t)ql); d += ql;
                *d++ = ':'; d = put_i64_fwd(d, q1, n1);
                *d++ = '-'; put_i64_fwd(d, q2, n2);
                slot[7][k] = o;
            }
            o = PyList_GET_ITEM(strand_objs, r.rec); Py_INCREF(o); slot[8][k] = o;
            o = PyList_GET_ITEM(ai_objs, r.rec); Py_INCREF(o); slot[10][k] = o;
            if (!(slot[11][k] = PyLong_FromLong(r.left_shift))) goto done;
            if (!(slot[12][k] = hom_str(r.hom_ref_l, r.hom_ref_r))) goto done;
            if (!(slot[13][k] = hom_str(r.hom_tig_l, r.hom_tig_r))) goto done;
            {   /* SEQ: DEL = reference[pos : pos+n] (unshifted); INS = forward contig[qry_pos : qry_end], reverse-complemented for
                 * minus-strand records (equals the slice of the reference-oriented contig, cigarcall.py:160,236) */
                Py_ssize_t w = is_del ? (Py_ssize_t)rid[r.rec] : n_ref + (Py_ssize_t)qid[r.rec];
                int64_t s0 = is_del ? r.pos : r.qry_pos, l = r.svlen;
                if (w < 0 || w >= nd_seq || s0 < 0 || l < 0 || s0 + l > bufs[w].len) { PyErr_SetString(PyExc_IndexError, "indel_frame: SEQ range outside sequence"); goto done; }
                const uint8_t *src = (const uint8_t *)bufs[w].buf + s0;
                if (!is_del && rv[r.rec]) {
                    if (l > cap) { PyMem_Free(scratch); cap = l * 2 + 64; scratch = (char *)PyMem_Malloc((size_t)cap); if (!scratch) { cap = 0; PyErr_NoMemory(); goto done; } }
                    for (int64_t j = 0; j < l; j++) scratch[j] = (char)ct[src[l - 1 - j]];
                    o = l == 1 ? one_char((unsigned char)scratch[0]) : PyUnicode_DecodeLatin1(scratch, l, NULL);
                } else o = l == 1 ? one_char(src[0]) : PyUnicode_DecodeLatin1((const char *)src, l, NULL);
                if (!o) goto done;
                if (l == 1) Py_INCREF(o);
                slot[15][k] = o;
            }
        }
    }
    result = PyTuple_New(INDEL_NCOL);
    if (result) for (int c = 0; c < INDEL_NCOL; c++) { PyTuple_SET_ITEM(result, c, cols[c]); cols[c] = NULL; }
done:
    for (int c = 0; c < INDEL_NCOL; c++) Py_XDECREF(cols[c]);
    PyMem_Free(scratch);
    if (bufs) { for (Py_ssize_t j = 0; j < got; j++) PyBuffer_Release(&bufs[j]); PyMem_Free(bufs); }
    strtab_free(&tc); strtab_free(&tq);
    PyBuffer_Release(&rows); PyBuffer_Release(&order); PyBuffer_Release(&ref_id); PyBuffer_Release(&qry_id); PyBuffer_Release(&rev); PyBuffer_Release(&comp);
    return result;
}

/* ------------------------------------------------------------------------------------------------
 * TSV writers: the text `DataFrame.to_csv(sep='\t', index=False)` would produce for the call tables (plus a FILTER
 * column), straight from the device rows -- for callers that write the table to a file at once, as `rule call_cigar` does
 * (rules/call.snakefile:813-846), without creating a single per-row Python object. No header line; fields are written
 * unquoted (the caller falls back to pandas when a name would need quoting).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { char *p; size_t len, cap; } obuf_t;

static int obuf_reserve(obuf_t *b, size_t more)
{
    if (b->len + more <= b->cap) return 0;
    size_t cap = b->cap ? b->cap : (1u << 20);
    while (cap < b->len + more) cap += cap / 2 + (1u << 20);
    char *q = (char *)realloc(b->p, cap);
    if (!q) { PyErr_NoMemory(); return -1; }
    b->p = q; b->cap = cap;
    return 0;
}

static inline void ob_mem(obuf_t *b, const char *s, size_t n) { memcpy(b->p + b->len, s, n); b->len += n; }
static inline void ob_ch(obuf_t *b, char c) { b->p[b->len++] = c; }
static inline void ob_i64(obuf_t *b, int64_t v) { int nd = ndig_i64(v); put_i64_fwd(b->p + b->len, v, nd); b->len += (size_t)nd; }

static size_t strtab_max(const strtab_t *t)
{
    size_t m = 0;
    for (Py_ssize_t i = 0; i < t->n; i++) if ((size_t)t->len[i] > m) m = (size_t)t->len[i];
    return m;
}

/* snv_tsv(rows, order, ids|None, chrom_strs, qry_strs, strand_strs, ai_strs, ref_id, qry_id, rev, seqs, n_ref, comp, hap, source, pass)
 *   as snv_frame, with per-record *strings* for every per-record column, `pass`: uint8 per emission row (1 = PASS, 0 = TRIM),
 *   or None for no FILTER column -> bytes */
static PyObject *py_snv_tsv(PyObject *self, PyObject *args)
{
    Py_buffer rows, order, ref_id, qry_id, rev, comp;
    PyObject *ids, *chrom_strs, *qry_strs, *strand_strs, *ai_strs, *seqs, *pass_obj;
    const char *hap, *source, *header; Py_ssize_t hap_n, source_n, header_n, n_ref;
    if (!PyArg_ParseTuple(args, "y*y*OOOOOy*y*y*O!ny*s#s#Oy#", &rows, &order, &ids, &chrom_strs, &qry_strs, &strand_strs, &ai_strs,
                          &ref_id, &qry_id, &rev, &PyList_Type, &seqs, &n_ref, &comp, &hap, &hap_n, &source, &source_n, &pass_obj, &header, &header_n))
        return NULL;
    PyObject *result = NULL;
    strtab_t tc = {0}, tq = {0}, ts = {0}, ta = {0};
    PyObject **id_src = NULL;
    Py_buffer pass; int has_pass = 0;
    obuf_t ob = {0};
    Py_ssize_t n = order.len / 8, n_rows = rows.len / (Py_ssize_t)sizeof(snv_row_t), nd_seq = PyList_GET_SIZE(seqs), got = 0;
    Py_buffer *bufs = (Py_buffer *)PyMem_Calloc((size_t)nd_seq + 1, sizeof(Py_buffer));
    if (!bufs) { PyErr_NoMemory(); goto done; }
    if (pass_obj != Py_None) { if (PyObject_GetBuffer(pass_obj, &pass, PyBUF_SIMPLE) < 0) goto done; has_pass = 1; }
    if (n > n_rows || comp.len < 256 || (has_pass && pass.len < n_rows)) { PyErr_SetString(PyExc_ValueError, "snv_tsv: buffer sizes disagree"); goto done; }
    if (strtab_load(&tc, chrom_strs, "chrom_strs") < 0 || strtab_load(&tq, qry_strs, "qry_strs") < 0 || strtab_load(&ts, strand_strs, "strand_strs") < 0 ||
        strtab_load(&ta, ai_strs, "ai_strs") < 0) goto done;
    Py_ssize_t n_rec = tc.n;
    if (tq.n < n_rec || ts.n < n_rec || ta.n < n_rec || ref_id.len < n_rec * 4 || qry_id.len < n_rec * 4 || rev.len < n_rec) {
        PyErr_SetString(PyExc_ValueError, "snv_tsv: per-record arrays disagree"); goto done;
    }
    if (ids != Py_None) {
        if (!PyArray_Check(ids) || PyArray_TYPE((PyArrayObject *)ids) != NPY_OBJECT || PyArray_NDIM((PyArrayObject *)ids) != 1 ||
            PyArray_DIM((PyArrayObject *)ids, 0) < n_rows || !PyArray_IS_C_CONTIGUOUS((PyArrayObject *)ids)) {
            PyErr_SetString(PyExc_TypeError, "snv_tsv: ids must be a contiguous 1-D object array with one entry per row"); goto done;
        }
        id_src = (PyObject **)PyArray_DATA((PyArrayObject *)ids);
    }
    for (; got < nd_seq; got++)
        if (PyObject_GetBuffer(PyList_GET_ITEM(seqs, got), &bufs[got], PyBUF_SIMPLE) < 0) goto done;
    {
        const snv_row_t *R = (const snv_row_t *)rows.buf;
        const int64_t *ord = (const int64_t *)order.buf;
        const int32_t *rid = (const int32_t *)ref_id.buf, *qid = (const int32_t *)qry_id.buf;
        const uint8_t *rv = (const uint8_t *)rev.buf, *ct = (const uint8_t *)comp.buf;
        if (obuf_reserve(&ob, (size_t)header_n + (size_t)n * 96) < 0) goto done;   /* header line first; a first guess at the size */
        ob_mem(&ob, header, (size_t)header_n);
        const size_t fixed = 2 * strtab_max(&tc) + strtab_max(&tq) + strtab_max(&ts) + strtab_max(&ta) + (size_t)hap_n + (size_t)source_n + 160;
        for (Py_ssize_t k = 0; k < n; k++) {
            int64_t i = ord[k];
            prefetch_snv(R, ord, k, n, n_rows, n_rec, rid, qid, bufs, n_ref, nd_seq);
            if (i < 0 || i >= n_rows) { PyErr_SetString(PyExc_IndexError, "snv_tsv: order entry out of range"); goto done; }
            const snv_row_t r = R[i];
            if (r.rec < 0 || r.rec >= n_rec) { PyErr_SetString(PyExc_IndexError, "snv_tsv: record index out of range"); goto done; }
            const Py_ssize_t wr = rid[r.rec], wq = n_ref + (Py_ssize_t)qid[r.rec];
            if (wr < 0 || wr >= n_ref || wq < n_ref || wq >= nd_seq || r.pos_ref < 0 || r.pos_ref >= bufs[wr].len || r.qry_pos < 0 || r.qry_pos >= bufs[wq].len) {
                PyErr_SetString(PyExc_IndexError, "snv_tsv: position outside its sequence"); goto done;
            }
            const unsigned rb = ((const uint8_t *)bufs[wr].buf)[r.pos_ref];
            unsigned ab = ((const uint8_t *)bufs[wq].buf)[r.qry_pos];
            if (rv[r.rec]) ab = ct[ab];
            size_t need = fixed;
            const char *idp = NULL; Py_ssize_t idn = 0;
            if (id_src) { idp = PyUnicode_AsUTF8AndSize(id_src[i], &idn); if (!idp) goto done; need += (size_t)idn; }
            if (obuf_reserve(&ob, need) < 0) goto done;
            const int64_t p1 = (int64_t)r.pos_ref + 1, q1 = (int64_t)r.qry_pos + 1;
            ob_mem(&ob, tc.s[r.rec], (size_t)tc.len[r.rec]); ob_ch(&ob, '\t');
            ob_i64(&ob, r.pos_ref); ob_ch(&ob, '\t'); ob_i64(&ob, p1); ob_ch(&ob, '\t');
            if (idp) ob_mem(&ob, idp, (size_t)idn);
            else {
                ob_mem(&ob, tc.s[r.rec], (size_t)tc.len[r.rec]); ob_ch(&ob, '-'); ob_i64(&ob, p1); ob_mem(&ob, "-SNV-", 5);
                ob_ch(&ob, (char)((rb >= 'a' && rb <= 'z') ? rb - 32 : rb)); ob_ch(&ob, (char)((ab >= 'a' && ab <= 'z') ? ab - 32 : ab));
            }
            ob_mem(&ob, "\tSNV\t1\t", 7); ob_ch(&ob, (char)rb); ob_ch(&ob, '\t'); ob_ch(&ob, (char)ab); ob_ch(&ob, '\t');
            ob_mem(&ob, hap, (size_t)hap_n); ob_ch(&ob, '\t');
            ob_mem(&ob, tq.s[r.rec], (size_t)tq.len[r.rec]); ob_ch(&ob, ':'); ob_i64(&ob, q1); ob_ch(&ob, '-'); ob_i64(&ob, q1); ob_ch(&ob, '\t');
            ob_mem(&ob, ts.s[r.rec], (size_t)ts.len[r.rec]); ob_mem(&ob, "\t0\t", 3);
            ob_mem(&ob, ta.s[r.rec], (size_t)ta.len[r.rec]); ob_ch(&ob, '\t');
            ob_mem(&ob, source, (size_t)source_n);
            if (has_pass) ob_mem(&ob, ((const uint8_t *)pass.buf)[i] ? "\tPASS" : "\tTRIM", 5);
            ob_ch(&ob, '\n');
        }
    }
    result = PyBytes_FromStringAndSize(ob.p ? ob.p : "", (Py_ssize_t)ob.len);
done:
    free(ob.p);
    if (has_pass) PyBuffer_Release(&pass);
    if (bufs) { for (Py_ssize_t j = 0; j < got; j++) PyBuffer_Release(&bufs[j]); PyMem_Free(bufs); }
    strtab_free(&tc); strtab_free(&tq); strtab_free(&ts); strtab_free(&ta);
    PyBuffer_Release(&rows); PyBuffer_Release(&order); PyBuffer_Release(&ref_id); PyBuffer_Release(&qry_id); PyBuffer_Release(&rev); PyBuffer_Release(&comp);
    return result;
}

/* indel_tsv(rows, order, ids|None, chrom_strs, qry_strs, strand_strs, ai_strs, ref_id, qry_id, rev, seqs, n_ref, comp, hap, source, pass) -> bytes */
static PyObject *py_indel_tsv(PyObject *self, PyObject *args)
{
    Py_buffer rows, order, ref_id, qry_id, rev, comp;
    PyObject *ids, *chrom_strs, *qry_strs, *strand_strs, *ai_strs, *seqs, *pass_obj;
    const char *hap, *source, *header; Py_ssize_t hap_n, source_n, header_n, n_ref;
    if (!PyArg_ParseTuple(args, "y*y*OOOOOy*y*y*O!ny*s#s#Oy#", &rows, &order, &ids, &chrom_strs, &qry_strs, &strand_strs, &ai_strs,
                          &ref_id, &qry_id, &rev, &PyList_Type, &seqs, &n_ref, &comp, &hap, &hap_n, &source, &source_n, &pass_obj, &header, &header_n))
        return NULL;
    PyObject *result = NULL;
    strtab_t tc = {0}, tq = {0}, ts = {0}, ta = {0};
    PyObject **id_src = NULL;
    Py_buffer pass; int has_pass = 0;
    obuf_t ob = {0};
    Py_ssize_t n = order.len / 8, n_rows = rows.len / (Py_ssize_t)sizeof(indel_row_t), nd_seq = PyList_GET_SIZE(seqs), got = 0;
    Py_buffer *bufs = (Py_buffer *)PyMem_Calloc((size_t)nd_seq + 1, sizeof(Py_buffer));
    if (!bufs) { PyErr_NoMemory(); goto done; }
    if (pass_obj != Py_None) { if (PyObject_GetBuffer(pass_obj, &pass, PyBUF_SIMPLE) < 0) goto done; has_pass = 1; }
    if (n > n_rows || comp.len < 256 || (has_pass && pass.len < n_rows)) { PyErr_SetString(PyExc_ValueError, "indel_tsv: buffer sizes disagree"); goto done; }
    if (strtab_load(&tc, chrom_strs, "chrom_strs") < 0 || strtab_load(&tq, qry_strs, "qry_strs") < 0 || strtab_load(&ts, strand_strs, "strand_strs") < 0 ||
        strtab_load(&ta, ai_strs, "ai_strs") < 0) goto done;
    Py_ssize_t n_rec = tc.n;
    if (tq.n < n_rec || ts.n < n_rec || ta.n < n_rec || ref_id.len < n_rec * 4 || qry_id.len < n_rec * 4 || rev.len < n_rec) {
        PyErr_SetString(PyExc_ValueError, "indel_tsv: per-record arrays disagree"); goto done;
    }
    if (ids != Py_None) {
        if (!PyArray_Check(ids) || PyArray_TYPE((PyArrayObject *)ids) != NPY_OBJECT || PyArray_NDIM((PyArrayObject *)ids) != 1 ||
            PyArray_DIM((PyArrayObject *)ids, 0) < n_rows || !PyArray_IS_C_CONTIGUOUS((PyArrayObject *)ids)) {
            PyErr_SetString(PyExc_TypeError, "indel_tsv: ids must be a contiguous 1-D object array with one entry per row"); goto done;
        }
        id_src = (PyObject **)PyArray_DATA((PyArrayObject *)ids);
    }
    for (; got < nd_seq; got++)
        if (PyObject_GetBuffer(PyList_GET_ITEM(seqs, got), &bufs[got], PyBUF_SIMPLE) < 0) goto done;
    {
        const indel_row_t *R = (const indel_row_t *)rows.buf;
        const int64_t *ord = (const int64_t *)order.buf;
        const int32_t *rid = (const int32_t *)ref_id.buf, *qid = (const int32_t *)qry_id.buf;
        const uint8_t *rv = (const uint8_t *)rev.buf, *ct = (const uint8_t *)comp.buf;
        if (obuf_reserve(&ob, (size_t)header_n + (size_t)n * 128) < 0) goto done;
        ob_mem(&ob, header, (size_t)header_n);
        const size_t fixed = 2 * strtab_max(&tc) + strtab_max(&tq) + strtab_max(&ts) + strtab_max(&ta) + (size_t)hap_n + (size_t)source_n + 400;
        for (Py_ssize_t k = 0; k < n; k++) {
            int64_t i = ord[k];
            if (i < 0 || i >= n_rows) { PyErr_SetString(PyExc_IndexError, "indel_tsv: order entry out of range"); goto done; }
            const indel_row_t r = R[i];
            if (r.rec < 0 || r.rec >= n_rec) { PyErr_SetString(PyExc_IndexError, "indel_tsv: record index out of range"); goto done; }
            const int is_del = r.svtype == 1;
            const Py_ssize_t w = is_del ? (Py_ssize_t)rid[r.rec] : n_ref + (Py_ssize_t)qid[r.rec];
            const int64_t s0 = is_del ? r.pos : r.qry_pos, l = r.svlen;
            if (w < 0 || w >= nd_seq || s0 < 0 || l < 0 || s0 + l > bufs[w].len) { PyErr_SetString(PyExc_IndexError, "indel_tsv: SEQ range outside sequence"); goto done; }
            size_t need = fixed + (size_t)l;
            const char *idp = NULL; Py_ssize_t idn = 0;
            if (id_src) { idp = PyUnicode_AsUTF8AndSize(id_src[i], &idn); if (!idp) goto done; need += (size_t)idn; }
            if (obuf_reserve(&ob, need) < 0) goto done;
            const int64_t p1 = (int64_t)r.pos + 1, q1 = (int64_t)r.qry_pos + 1, q2 = is_del ? q1 : (int64_t)r.qry_end;
            ob_mem(&ob, tc.s[r.rec], (size_t)tc.len[r.rec]); ob_ch(&ob, '\t');
            ob_i64(&ob, r.pos); ob_ch(&ob, '\t'); ob_i64(&ob, r.end); ob_ch(&ob, '\t');
            if (idp) ob_mem(&ob, idp, (size_t)idn);
            else { ob_mem(&ob, tc.s[r.rec], (size_t)tc.len[r.rec]); ob_ch(&ob, '-'); ob_i64(&ob, p1); ob_mem(&ob, is_del ? "-DEL-" : "-INS-", 5); ob_i64(&ob, r.svlen); }
            ob_mem(&ob, is_del ? "\tDEL\t" : "\tINS\t", 5); ob_i64(&ob, r.svlen); ob_ch(&ob, '\t');
            ob_mem(&ob, hap, (size_t)hap_n); ob_ch(&ob, '\t');
            ob_mem(&ob, tq.s[r.rec], (size_t)tq.len[r.rec]); ob_ch(&ob, ':'); ob_i64(&ob, q1); ob_ch(&ob, '-'); ob_i64(&ob, q2); ob_ch(&ob, '\t');
            ob_mem(&ob, ts.s[r.rec], (size_t)ts.len[r.rec]); ob_mem(&ob, "\t0\t", 3);
            ob_mem(&ob, ta.s[r.rec], (size_t)ta.len[r.rec]); ob_ch(&ob, '\t');
            ob_i64(&ob, r.left_shift); ob_ch(&ob, '\t');
            ob_i64(&ob, r.hom_ref_l); ob_ch(&ob, ','); ob_i64(&ob, r.hom_ref_r); ob_ch(&ob, '\t');
            ob_i64(&ob, r.hom_tig_l); ob_ch(&ob, ','); ob_i64(&ob, r.hom_tig_r); ob_ch(&ob, '\t');
            ob_mem(&ob, source, (size_t)source_n); ob_ch(&ob, '\t');
            {
                const uint8_t *src = (const uint8_t *)bufs[w].buf + s0;
                char *d = ob.p + ob.len;
                if (!is_del && rv[r.rec]) for (int64_t j = 0; j < l; j++) d[j] = (char)ct[src[l - 1 - j]];
                else memcpy(d, src, (size_t)l);
                ob.len += (size_t)l;
            }
            if (has_pass) ob_mem(&ob, ((const uint8_t *)pass.buf)[i] ? "\tPASS" : "\tTRIM", 5);
            ob_ch(&ob, '\n');
        }
    }
    result = PyBytes_FromStringAndSize(ob.p ? ob.p : "", (Py_ssize_t)ob.len);
done:
    free(ob.p);
    if (has_pass) PyBuffer_Release(&pass);
    if (bufs) { for (Py_ssize_t j = 0; j < got; j++) PyBuffer_Release(&bufs[j]); PyMem_Free(bufs); }
    strtab_free(&tc); strtab_free(&tq); strtab_free(&ts); strtab_free(&ta);
    PyBuffer_Release(&rows); PyBuffer_Release(&order); PyBuffer_Release(&ref_id); PyBuffer_Release(&qry_id); PyBuffer_Release(&rev); PyBuffer_Release(&comp);
    return result;
}

/* ------------------------------------------------------------------------------------------------
 * Arena recycling. A 2 M-row result is ~8 M small str / int objects = several hundred MB of pymalloc arenas. CPython maps
 * every arena fresh (mmap, 1 MiB) and unmaps it as soon as it is empty, so each call pays the first-touch page faults
 * for all of it again (measured: ~0.4 ns per byte, a third of the frame-building time). keep_arenas(max_mb) installs an
 * arena allocator (PyObject_SetArenaAllocator, the documented hook) that keeps up to max_mb of released arenas and hands them
 * out again; anything beyond goes back to the OS as before. Arenas need not be zeroed (pymalloc initialises the pools it carves).
 * ------------------------------------------------------------------------------------------------ */
static struct {
    void **ptr;
    size_t n, cap, size;
    int installed;
    PyObjectArenaAllocator prev;   /* the allocator that was installed before: arenas are obtained from it and returned to it */
} g_arena;

static void *arena_alloc(void *ctx, size_t size)
{
    if (g_arena.n > 0 && size == g_arena.size) return g_arena.ptr[--g_arena.n];
    return g_arena.prev.alloc(g_arena.prev.ctx, size);
}

static void arena_free(void *ctx, void *ptr, size_t size)
{
    if (g_arena.n < g_arena.cap && (g_arena.size == 0 || g_arena.size == size)) {
        g_arena.size = size;
        g_arena.ptr[g_arena.n++] = ptr;
        return;
    }
    g_arena.prev.free(g_arena.prev.ctx, ptr, size);
}

/* Opt-in (PAVGPU_TUNE_ALLOC=1, pavlib/cigarcall.py). Chains to the previous arena allocator, so it makes no assumption about how
 * arenas are mapped; refused on free-threaded builds (the cache above relies on the GIL). */
static PyObject *py_keep_arenas(PyObject *self, PyObject *arg)
{
    long max_mb = PyLong_AsLong(arg);
    if (max_mb == -1 && PyErr_Occurred()) return NULL;
#ifdef Py_GIL_DISABLED
    Py_RETURN_FALSE;
#endif
    if (g_arena.installed || max_mb <= 0) Py_RETURN_FALSE;
    g_arena.cap = (size_t)max_mb;   /* arenas are 1 MiB on 64-bit CPython >= 3.10 (256 KiB before: the cap is then a quarter) */
    g_arena.ptr = (void **)malloc(g_arena.cap * sizeof(void *));
    if (!g_arena.ptr) return PyErr_NoMemory();
    PyObject_GetArenaAllocator(&g_arena.prev);
    if (!g_arena.prev.alloc || !g_arena.prev.free) { free(g_arena.ptr); g_arena.ptr = NULL; Py_RETURN_FALSE; }
    PyObjectArenaAllocator a = {NULL, arena_alloc, arena_free};
    PyObject_SetArenaAllocator(&a);
    g_arena.installed = 1;
    Py_RETURN_TRUE;
}

static PyMethodDef methods[] = {
    {"format", py_format, METH_VARARGS, "format(n, parts) -> list of str"},
    {"ints", py_ints, METH_O, "ints(int64 buffer) -> list of int"},
    {"slices", py_slices, METH_VARARGS, "slices(data list, which, start, length, rc, comp) -> list of str"},
    {"snv_tsv", py_snv_tsv, METH_VARARGS, "TSV text of df_snv (+ FILTER) without building the frame"},
    {"indel_tsv", py_indel_tsv, METH_VARARGS, "TSV text of df_insdel (+ FILTER) without building the frame"},
    {"keep_arenas", py_keep_arenas, METH_O, "keep_arenas(max_mb): recycle up to max_mb of released pymalloc arenas"},
    {"snv_frame", py_snv_frame, METH_VARARGS, "all 14 columns of df_snv in final row order"},
    {"indel_frame", py_indel_frame, METH_VARARGS, "all 16 columns of df_insdel in final row order"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pyrows", "row formatting helpers", -1, methods};

PyMODINIT_FUNC PyInit__pyrows(void)
{
    import_array();
    return PyModule_Create(&moddef);
}
