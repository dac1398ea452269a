/*
 * Thin CPython wrapper that exposes the UNMODIFIED reference ray tracer
 * (/root/reference/src/cradon.c, compiled from where it lies) as
 *     ref_cradon.trace(Tx, Rx, nx, ny) -> (data f64, indices i64, indptr i64)
 * so that the CSR integers can be diffed directly (src/radon.pyx only exposes
 * the sinogram).  Test infrastructure only; output goes to oracle/_ref/.
 */
#include "cradon.h"

static PyObject *trace(PyObject *self, PyObject *args)
{
    PyObject *otx, *orx;
    Py_ssize_t nx, ny;
    if (!PyArg_ParseTuple(args, "OOnn", &otx, &orx, &nx, &ny)) return NULL;
    PyArrayObject *tx = (PyArrayObject *)PyArray_FROM_OTF(otx, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
    PyArrayObject *rx = (PyArrayObject *)PyArray_FROM_OTF(orx, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
    if (!tx || !rx) { Py_XDECREF(tx); Py_XDECREF(rx); return NULL; }
    PyObject *L = PyTuple_New(3);
    radon2d((const double *)PyArray_DATA(tx), (const double *)PyArray_DATA(rx),
            (size_t)PyArray_DIM(tx, 0), (size_t)nx, (size_t)ny, L);
    Py_DECREF(tx);
    Py_DECREF(rx);
    return L;
}

static PyMethodDef methods[] = {
    {"trace", trace, METH_VARARGS, "reference cradon.radon2d -> CSR tuple"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "ref_cradon", NULL, -1, methods};

PyMODINIT_FUNC PyInit_ref_cradon(void)
{
    import_array();
    return PyModule_Create(&moddef);
}
