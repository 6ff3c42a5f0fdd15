/* Stand-in for libipopt's C interface (IpStdCInterface.h) used ONLY to test the
 * ctypes binding of colloc_fem_code_b200/nlp.py where no IPOPT is installed.
 * "Solving" means: query the structure, evaluate every callback once at the
 * start point with lambda_i = 0.5 + 0.01*i and obj_factor = 0.75, and report
 * what the callbacks returned through the output arguments. */
#include <stdlib.h>
#include <string.h>

typedef double Number; typedef int Index; typedef int Int; typedef int Bool;
typedef void* UserDataPtr;
typedef Bool (*Eval_F_CB)(Index, Number*, Bool, Number*, UserDataPtr);
typedef Bool (*Eval_Grad_F_CB)(Index, Number*, Bool, Number*, UserDataPtr);
typedef Bool (*Eval_G_CB)(Index, Number*, Bool, Index, Number*, UserDataPtr);
typedef Bool (*Eval_Jac_G_CB)(Index, Number*, Bool, Index, Index, Index*, Index*, Number*, UserDataPtr);
typedef Bool (*Eval_H_CB)(Index, Number*, Bool, Number, Index, Number*, Bool, Index, Index*, Index*, Number*, UserDataPtr);

struct Problem {
    Index n, m, nele_jac, nele_hess, index_style;
    Eval_F_CB f; Eval_G_CB g; Eval_Grad_F_CB grad; Eval_Jac_G_CB jac; Eval_H_CB h;
    double tol; int max_iter; char linear_solver[64]; double obj_scaling;
    int have_scaling; double x_scale0, g_scale0;
};

struct Problem* CreateIpoptProblem(Index n, Number* x_L, Number* x_U, Index m, Number* g_L, Number* g_U,
                                   Index nele_jac, Index nele_hess, Index index_style,
                                   Eval_F_CB f, Eval_G_CB g, Eval_Grad_F_CB grad, Eval_Jac_G_CB jac, Eval_H_CB h)
{
    struct Problem* p = calloc(1, sizeof *p);
    (void)x_L; (void)x_U; (void)g_L; (void)g_U;
    p->n = n; p->m = m; p->nele_jac = nele_jac; p->nele_hess = nele_hess; p->index_style = index_style;
    p->f = f; p->g = g; p->grad = grad; p->jac = jac; p->h = h;
    return p;
}
void FreeIpoptProblem(struct Problem* p) { free(p); }
Bool AddIpoptStrOption(struct Problem* p, char* k, char* v)
{ if (!strcmp(k, "linear_solver")) strncpy(p->linear_solver, v, 63); return 1; }
Bool AddIpoptNumOption(struct Problem* p, char* k, Number v) { if (!strcmp(k, "tol")) p->tol = v; return 1; }
Bool AddIpoptIntOption(struct Problem* p, char* k, Int v) { if (!strcmp(k, "max_iter")) p->max_iter = v; return 1; }
Bool SetIpoptProblemScaling(struct Problem* p, Number obj, Number* xs, Number* gs)
{ p->obj_scaling = obj; p->have_scaling = 1; p->x_scale0 = xs ? xs[0] : 0; p->g_scale0 = gs ? gs[0] : 0; return 1; }

Int IpoptSolve(struct Problem* p, Number* x, Number* g, Number* obj_val, Number* mult_g,
               Number* mult_x_L, Number* mult_x_U, UserDataPtr ud)
{
    Index i;
    Index* jr = malloc(sizeof(Index) * (p->nele_jac + 1)); Index* jc = malloc(sizeof(Index) * (p->nele_jac + 1));
    Index* hr = malloc(sizeof(Index) * (p->nele_hess + 1)); Index* hc = malloc(sizeof(Index) * (p->nele_hess + 1));
    Number* jv = malloc(sizeof(Number) * (p->nele_jac + 1)); Number* hv = malloc(sizeof(Number) * (p->nele_hess + 1));
    Number* grad = malloc(sizeof(Number) * (p->n + 1));
    int ok = 1;
    double sj = 0, sh = 0, sg = 0, sidx = 0;
    for (i = 0; i < p->m; ++i) mult_g[i] = 0.5 + 0.01 * i;
    ok &= p->jac(p->n, x, 0, p->m, p->nele_jac, jr, jc, NULL, ud);       /* structure */
    ok &= p->h(p->n, x, 0, 1.0, p->m, mult_g, 0, p->nele_hess, hr, hc, NULL, ud);
    ok &= p->f(p->n, x, 1, obj_val, ud);
    ok &= p->grad(p->n, x, 0, grad, ud);
    ok &= p->g(p->n, x, 0, p->m, g, ud);
    ok &= p->jac(p->n, x, 0, p->m, p->nele_jac, NULL, NULL, jv, ud);
    ok &= p->h(p->n, x, 0, 0.75, p->m, mult_g, 1, p->nele_hess, NULL, NULL, hv, ud);
    for (i = 0; i < p->nele_jac; ++i) { sj += jv[i]; sidx += jr[i] + 2.0 * jc[i]; }
    for (i = 0; i < p->nele_hess; ++i) { sh += hv[i]; sidx += 3.0 * hr[i] + 5.0 * hc[i]; }
    for (i = 0; i < p->n; ++i) sg += grad[i];
    for (i = 0; i < p->n; ++i) { mult_x_L[i] = 0; mult_x_U[i] = 0; }
    mult_x_L[0] = sj; mult_x_L[1] = sh; mult_x_L[2] = sg; mult_x_L[3] = sidx;
    mult_x_U[0] = p->tol; mult_x_U[1] = p->max_iter; mult_x_U[2] = p->obj_scaling;
    mult_x_U[3] = p->have_scaling ? p->x_scale0 + 10 * p->g_scale0 : -1;
    mult_x_U[4] = !strcmp(p->linear_solver, "ma57");
    free(jr); free(jc); free(hr); free(hc); free(jv); free(hv); free(grad);
    return ok ? 0 : -13;   /* 0 = Solve_Succeeded, -13 = Invalid_Number_Detected */
}
