/* pylists.c -- host-side helper (CPython C API): packed interval arrays -> the reference's return type.
 *
 * NeuralSemiCRFInterval.decode() returns List[List[Tuple[int, int]]] (reference CRF/NeuralSemiCRFInterval.py:100-104);
 * at T=2048, N=88 that is ~1.6e5 tuples, and building them with zip()/tolist() costs more than the PCIe upload of the
 * score tensor (bench.py e2e.breakdown).  This does it in one C loop over the device's packed result.
 *
 *   pairs_to_lists(pairs: buffer int32 [N][stride][2], counts: buffer int32 [N], stride: int) -> list[list[tuple[int,int]]]
 *
 * Positions are small non-negative integers that repeat across tracks: the int objects for 0 .. kIntCache-1 are created
 * once and shared (one allocation per interval -- the tuple -- instead of three, and nothing but reference counts to
 * touch when the result is freed), exactly like CPython's own cache for -5 .. 256.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

enum { kIntCache = 1 << 16 };
static PyObject *int_cache[kIntCache];   /* lazily filled, lives as long as the module */

static inline PyObject *position(long v) {
    if (v >= 0 && v < kIntCache) {
        PyObject *o = int_cache[v];
        if (!o) {
            o = PyLong_FromLong(v);
            if (!o) return NULL;
            int_cache[v] = o;   /* the cache keeps this reference */
        }
        Py_INCREF(o);
        return o;
    }
    return PyLong_FromLong(v);
}

static PyObject *pairs_to_lists(PyObject *self, PyObject *args) {
    Py_buffer pb, cb;
    Py_ssize_t stride;
    if (!PyArg_ParseTuple(args, "y*y*n", &pb, &cb, &stride)) return NULL;
    const Py_ssize_t n = cb.len / (Py_ssize_t)sizeof(int32_t);
    const int32_t *pairs = (const int32_t *)pb.buf, *counts = (const int32_t *)cb.buf;
    PyObject *out = NULL;
    if (stride < 0 || pb.len < n * stride * 2 * (Py_ssize_t)sizeof(int32_t)) {
        PyErr_SetString(PyExc_ValueError, "pairs buffer smaller than [len(counts)][stride][2] int32");
        goto done;
    }
    out = PyList_New(n);
    if (!out) goto done;
    for (Py_ssize_t t = 0; t < n; ++t) {
        Py_ssize_t c = counts[t];
        if (c < 0 || c > stride) {
            PyErr_SetString(PyExc_ValueError, "count out of range");
            Py_CLEAR(out);
            goto done;
        }
        PyObject *lst = PyList_New(c);
        if (!lst) {
            Py_CLEAR(out);
            goto done;
        }
        PyList_SET_ITEM(out, t, lst);
        const int32_t *p = pairs + t * stride * 2;
        for (Py_ssize_t i = 0; i < c; ++i) {
            PyObject *b = position(p[2 * i]), *e = position(p[2 * i + 1]);
            PyObject *tup = (b && e) ? PyTuple_New(2) : NULL;
            if (!tup) {
                Py_XDECREF(b);
                Py_XDECREF(e);
                Py_CLEAR(out);
                goto done;
            }
            PyTuple_SET_ITEM(tup, 0, b);
            PyTuple_SET_ITEM(tup, 1, e);
            PyList_SET_ITEM(lst, i, tup);
        }
    }
done:
    PyBuffer_Release(&pb);
    PyBuffer_Release(&cb);
    return out;
}

static PyMethodDef methods[] = {{"pairs_to_lists", pairs_to_lists, METH_VARARGS,
                                 "packed int32 pairs [N][stride][2] + counts [N] -> list of lists of (begin, end) tuples"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_tkb_pylists", NULL, -1, methods};
PyMODINIT_FUNC PyInit__tkb_pylists(void) { return PyModule_Create(&moduledef); }
