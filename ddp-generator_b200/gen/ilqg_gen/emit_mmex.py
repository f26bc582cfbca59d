"""Emit the single-evaluation mex gateway a reference build gets from iLQG_MMex.tem: `iLQG<Name>MMex.c`.

    out = iLQG<Name>MMex(x, u, params, mode, k, n_hor)

17 modes (iLQG_MMex.tem:81-226): 0 f, 1 L, 2 F, 3 Fx, 4 Fxx, 5 Lx, 6 Lu, 7 Lxx, 8 Luu, 9 Lxu, 10 fx, 11 fu, 12 fxx, 13 fuu, 14 fxu,
15 y (empty), 16 clamped u.  Everything is evaluated at ONE point: all auxiliary values and their derivatives first, unguarded
(`check_nan_inf_mode: false`, iLQG_MMex.tem:14), then the requested block, as FULL matrices (`tri_matrix_mode: false`,
genenerator_main.mac:41): Hessians column-major with both triangles, the second-order dynamics as A(c, j, r) = d2 f_r / d c d j
(print_jaco2, genenerator_main.mac:228-250).  Like iLQG_func.c this is reference-ABI problem code: it is what the checker
compiles (over oracle/mex_stub) to pin mex/iLQG_MMex_b200.c, the same gateway on the GPU library.

The template has no multiplier or penalty-weight inputs, so -- as with the reference -- only problems without
hfe / hfi / hle / hli have an MMex.
"""
from __future__ import annotations

from .emit_c import HDR, CNames
from .lin import Scope, render_operand, render_rhs

MODES = ("f", "L", "F", "Fx", "Fxx", "Lx", "Lu", "Lxx", "Luu", "Lxu", "fx", "fu", "fxx", "fuu", "fxu", "y", "clampU")


def mode_dims(m, mode):
    """MATLAB dimensions of the value a mode returns."""
    nx, nu = m.nx, m.nu
    return {0: (nx, 1), 1: (1, 1), 2: (1, 1), 3: (1, nx), 4: (nx, nx), 5: (1, nx), 6: (1, nu), 7: (nx, nx), 8: (nu, nu), 9: (nx, nu),
            10: (nx, nx), 11: (nx, nu), 12: (nx, nx, nx), 13: (nu, nu, nx), 14: (nx, nu, nx), 15: (0, 1), 16: (nu, 1)}[mode]


def _render(sc, names, tname, indent="            "):
    out = []
    for it in sc.items:
        if it[0] == "tmp":
            out.append(f"{indent}const double {it[1]} = {render_rhs(it[2], names)};")
        elif it[0] == "out":
            out.append(f"{indent}{tname(it[1])} = {render_operand(it[2], names)};")
        elif it[0] == "sincos":
            out.append(f"{indent}double {it[1]}, {it[2]}; dm_sincos({render_operand(it[3], names)}, &{it[1]}, &{it[2]});")
        elif it[0] == "raw":
            out.append(it[1](names, indent))
    return "\n".join(out)


def _utri(r, c):
    return (c * (c + 1)) // 2 + r if r <= c else (r * (r + 1)) // 2 + c


def emit_mmex_c(m) -> str:
    if any(m.n_mu.values()):
        raise ValueError("the MMex interface has no multiplier / penalty inputs (iLQG_MMex.tem): problems with hfe/hfi/hle/hli have none")
    nx, nu = m.nx, m.nu
    N = CNames(m, aux_prefix="aux_")
    o = [HDR % m.name]
    o.append('#include "mex.h"\n#ifndef  HAVE_OCTAVE\n#include "matrix.h"\n#endif\n\n#include <math.h>\n#include "dm_math.h"\n')
    o.append("typedef struct paramDesc {\n  char *name;\n  int size;\n  int is_var;\n} tParamDesc;\n")
    o.append(f"int n_params= {len(m.params)};\n")
    for i, d in enumerate(m.params):
        o.append(f'tParamDesc p_name{i + 1}= {{"{d.name}", {d.size}, 0}};')
    o.append("int n_vars= 0;\n")
    o.append("tParamDesc *paramdesc[]= {" + ", ".join(f"&p_name{i + 1}" for i in range(len(m.params))) + ("0" if not m.params else "") + "};\n")
    o.append(r"""void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    int mode, i, k, si, m_, n_;
    mwSize dims[3];   /* the template declares int dims[3]; mxCreateNumericArray takes mwSize */
    const mxArray *mxParam;
    const mxArray *mxParams;
    double t0, *L1, *L2, N, *fx, *fxx, *fu, *fuu, *fxu, limit;
    double *x;
    double *u;
    double **p;
    (void)limit; (void)L1; (void)L2; (void)fx; (void)fxx; (void)fu; (void)fuu; (void)fxu; (void)dims; (void)t0;

    if(nrhs!=6) { mexErrMsgTxt("wrong number of arguments (6 expected)"); return; }
    if(nlhs!=1) { mexErrMsgTxt("wrong number of return values (1 expected)"); return; }
    if(mxGetNumberOfElements(prhs[0])!=%(nx)d) { mexErrMsgTxt("wrong number of elements in x (%(nx)d expected)"); return; }
    if(mxGetNumberOfElements(prhs[1])!=%(nu)d) { mexErrMsgTxt("wrong number of elements in u (%(nu)d expected)"); return; }
    if(mxGetNumberOfElements(prhs[2])!=1) { mexErrMsgTxt("wrong number of elements in params (1 expected)"); return; }
    if(mxGetNumberOfElements(prhs[3])!=1) { mexErrMsgTxt("wrong number of elements in mode (1 expected)"); return; }
    if(mxGetNumberOfElements(prhs[4])!=1) { mexErrMsgTxt("wrong number of elements in k (1 expected)"); return; }
    if(mxGetNumberOfElements(prhs[5])!=1) { mexErrMsgTxt("wrong number of elements in n_hor (1 expected)"); return; }

    mode= (int)mxGetScalar(prhs[3]);
    k= (int)mxGetScalar(prhs[4])-1;
    N= mxGetScalar(prhs[5]);
    x= mxGetPr(prhs[0]);
    u= mxGetPr(prhs[1]);
    (void)k; (void)x; (void)u;

    mxParams= prhs[2];
    if(!mxIsStruct(mxParams)) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "Input 3 must be a struct.\n");
    }

    p= mxMalloc(n_params*sizeof(double *));
    for(i=0; i<n_params; i++) {
        si= (paramdesc[i]->size==-1)? N+1: paramdesc[i]->size;
        if((mxParam= mxGetField(mxParams, 0, paramdesc[i]->name))==NULL) {
            mxFree(p);
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Parameter name '%%s' is not member of parameters struct.\n", paramdesc[i]->name);
        }
        m_= mxGetM(mxParam);
        n_= mxGetN(mxParam);
        if(mxIsSparse(mxParam) || !mxIsDouble(mxParam) || (m_!=1 && n_!=1) || (m_*n_!=si)) {
            mxFree(p);
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Parameter name '%%s' must be a vector length %%d.\n", paramdesc[i]->name, si);
        }
        p[i]= mxGetPr(mxParam);
    }
""" % {"nx": nx, "nu": nu})
    # ---- all auxiliary values and their derivatives, unguarded (print_aux(); print_deriv();) ------------------------------------
    sc = Scope("a")
    decl = []
    for a in m.aux + m.daux:
        decl.append(f"aux_{a.name}")
        sc.out(("var", f"aux_{a.name}"), a.expr, False)
    if decl:
        o.append("    double " + ", ".join(decl) + ";")
        o.append("    " + " ".join(f"(void){d};" for d in decl))
    o.append(_render(sc, N, lambda t: t[1], indent="    "))

    def block(mode, head, entries, extra=""):
        """entries: list of (lvalue, expr)"""
        s = Scope(f"m{mode}_")
        for lv, e in entries:
            s.out(("var", lv), e, False)
        return f"        case {mode}: {{ /* {MODES[mode]} */\n{head}\n" + _render(s, N, lambda t: t[1]) + f"\n{extra}            break; }}\n"

    def create(mode, var=None):
        d = mode_dims(m, mode)
        if len(d) == 2:
            s = f"            plhs[0]= mxCreateDoubleMatrix({d[0]}, {d[1]}, mxREAL);"
        else:
            s = f"            dims[0]= {d[0]}; dims[1]= {d[1]}; dims[2]= {d[2]};\n            plhs[0]= mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);"
        if var:
            s += f"\n            {var}= mxGetPr(plhs[0]);"
        return s

    by = lambda entries: {e.idx: e.expr for e in entries}
    fxx_, fuu_, fxu_ = by(m.fxx), by(m.fuu), by(m.fxu)
    cxx_, cuu_, cxu_, Fcxx_ = by(m.cxx), by(m.cuu), by(m.cxu), by(m.Fcxx)
    nan_inf = "            if(mxIsNaN(t0)) t0= mxGetInf();\n            plhs[0]= mxCreateDoubleMatrix(1, 1, mxREAL);\n            (*mxGetPr(plhs[0]))= t0;\n"
    o.append("    switch(mode) {")
    o.append(block(0, create(0, "L1"), [(f"L1[{i}]", e) for i, e in enumerate(m.f)]))
    o.append(block(1, "", [("t0", m.L)], nan_inf))
    o.append(block(2, "", [("t0", m.F)], nan_inf))
    o.append(block(3, create(3, "L1"), [(f"L1[{e.idx}]", e.expr) for e in m.Fcx]))
    o.append(block(4, create(4, "L2"), [(f"L2[{r + c * nx}]", Fcxx_[_utri(r, c)]) for c in range(nx) for r in range(nx)]))
    o.append(block(5, create(5, "L1"), [(f"L1[{e.idx}]", e.expr) for e in m.cx]))
    o.append(block(6, create(6, "L1"), [(f"L1[{e.idx}]", e.expr) for e in m.cu]))
    o.append(block(7, create(7, "L2"), [(f"L2[{r + c * nx}]", cxx_[_utri(r, c)]) for c in range(nx) for r in range(nx)]))
    o.append(block(8, create(8, "L2"), [(f"L2[{r + c * nu}]", cuu_[_utri(r, c)]) for c in range(nu) for r in range(nu)]))
    o.append(block(9, create(9, "L2"), [(f"L2[{r + c * nx}]", cxu_[r + c * nx]) for c in range(nu) for r in range(nx)]))
    o.append(block(10, create(10, "fx"), [(f"fx[{e.idx}]", e.expr) for e in m.fx]))
    o.append(block(11, create(11, "fu"), [(f"fu[{e.idx}]", e.expr) for e in m.fu]))
    # second-order dynamics: value index = c + j * n_c + r * n_c * n_j  (r = component of f, outermost)
    o.append(block(12, create(12, "fxx"), [(f"fxx[{c + j * nx + r * nx * nx}]", fxx_[r * m.nqxx + _utri(c, j)])
                                          for r in range(nx) for j in range(nx) for c in range(nx)]))
    o.append(block(13, create(13, "fuu"), [(f"fuu[{c + j * nu + r * nu * nu}]", fuu_[r * m.nquu + _utri(c, j)])
                                          for r in range(nx) for j in range(nu) for c in range(nu)]))
    o.append(block(14, create(14, "fxu"), [(f"fxu[{c + j * nx + r * nx * nu}]", fxu_[r * m.nqxu + c + j * nx])
                                          for r in range(nx) for j in range(nu) for c in range(nx)]))
    o.append("        case 15: /* y */\n            plhs[0]= mxCreateDoubleMatrix(0, 1, mxREAL);\n            break;\n")
    sc = Scope("m16_")
    for rec in m.h:
        sc.out(("var", "limit"), rec["limit"], False)
        j, cmp_ = rec["input"], (">" if rec["sign"] > 0 else "<")
        sc.raw(lambda names, ind, j=j, cmp_=cmp_: f"{ind}if(u[{j}] {cmp_} limit) u[{j}]= limit;")
    o.append(f"        case 16: {{ /* clampU */\n            plhs[0]= mxCreateDoubleMatrix({nu}, 1, mxREAL);\n"
             f"            for(i= 0; i < {nu}; i++)\n                mxGetPr(plhs[0])[i]= u[i];\n            u= mxGetPr(plhs[0]);\n"
             + _render(sc, N, lambda t: t[1]) + "\n            break; }\n")
    o.append("    }\n\n    mxFree(p);\n}\n")
    return "\n".join(o)
