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

static PyMethodDef methods[] = {
    {"format", py_format, METH_VARARGS, "format(n, parts) -> list of str"},
    {"ints", py_ints, METH_O, "ints(int64 buffer) -> list of int"},
    {"slices", py_slices, METH_VARARGS, "slices(data list, which, start, length, rc, comp) -> list of str"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pyrows", "row formatting helpers", -1, methods};

PyMODINIT_FUNC PyInit__pyrows(void)
{
    import_array();
    return PyModule_Create(&moddef);
}
