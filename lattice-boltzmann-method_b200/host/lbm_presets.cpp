// lbm_presets.cpp — the boundary rule lists of the reference's drivers, written against the public
// C ABI only (lbm_bc_add), one lbm_bc_op per slice assignment and in the drivers' order so that
// later assignments win at corners exactly as they do there.
#include "../../include/lbm_b200.h"

namespace
{
struct Rules
{
  lbm_domain* d;
  int status = LBM_OK;
  explicit Rules(lbm_domain* dom) : d(dom) { status = lbm_bc_clear(d); }

  void add(const lbm_bc_op& op)
  {
    if (status == LBM_OK) status = lbm_bc_add(d, &op);
  }
  static lbm_bc_op base(int lattice, int xb, int xe, int yb, int ye)
  {
    lbm_bc_op op;
    lbm_bc_op_default(&op);
    op.lattice = lattice;
    op.x_begin = xb; op.x_end = xe; op.y_begin = yb; op.y_end = ye;
    return op;
  }
  // f_adve[region, dq] = coef * f_coll[region, sq]
  void local(int lattice, int xb, int xe, int yb, int ye, int dq, int sq, double coef = 1.0)
  {
    lbm_bc_op op = base(lattice, xb, xe, yb, ye);
    op.kind = LBM_BC_LINEAR; op.dst_q = dq; op.src_q = sq; op.coef = coef;
    add(op);
  }
  // f_adve[region, dq] = f_coll[(row|col) src, dq]
  void from(int lattice, int xb, int xe, int yb, int ye, int dq, int mode, int where)
  {
    lbm_bc_op op = base(lattice, xb, xe, yb, ye);
    op.kind = LBM_BC_LINEAR; op.dst_q = dq; op.src_q = dq; op.src_mode = mode; op.src_a = where;
    add(op);
  }
  int commit() { return status == LBM_OK ? lbm_bc_commit(d) : status; }
};

// rows of a single index i (negative from the end): [i, i+1) with -1 -> [-1, END)
inline int one_end(int i) { return i == -1 ? LBM_END : i + 1; }

void pressure_rows(Rules& r, double rho_in, double rho_out)
{
  // f_coll[0,:]  = feq(rho_in , u[-2,:]) + f_coll[-2,:] - f_equi[-2,:]
  lbm_bc_op op = Rules::base(0, 0, 1, 0, LBM_END);
  op.kind = LBM_BC_PRESSURE_PERIODIC; op.src_mode = LBM_SRC_ROW; op.src_a = -2; op.rho_bc = rho_in;
  r.add(op);
  // f_coll[-1,:] = feq(rho_out, u[ 1,:]) + f_coll[ 1,:] - f_equi[ 1,:]
  op = Rules::base(0, -1, LBM_END, 0, LBM_END);
  op.kind = LBM_BC_PRESSURE_PERIODIC; op.src_mode = LBM_SRC_ROW; op.src_a = 1; op.rho_bc = rho_out;
  r.add(op);
}

void bounce_back_columns(Rules& r)
{
  r.local(0, 0, LBM_END, -1, LBM_END, 4, 2);
  r.local(0, 0, LBM_END, -1, LBM_END, 7, 5);
  r.local(0, 0, LBM_END, -1, LBM_END, 8, 6);
  r.local(0, 0, LBM_END, 0, 1, 2, 4);
  r.local(0, 0, LBM_END, 0, 1, 5, 7);
  r.local(0, 0, LBM_END, 0, 1, 6, 8);
}

void specular_columns(Rules& r)
{
  r.local(0, 0, LBM_END, -1, LBM_END, 4, 2);
  r.local(0, 0, LBM_END, -1, LBM_END, 7, 6);
  r.local(0, 0, LBM_END, -1, LBM_END, 8, 5);
  r.local(0, 0, LBM_END, 0, 1, 2, 4);
  r.local(0, 0, LBM_END, 0, 1, 5, 8);
  r.local(0, 0, LBM_END, 0, 1, 6, 7);
}
}  // namespace

extern "C"
{

int lbm_preset_periodic(lbm_domain* d)
{
  Rules r(d);
  return r.commit();
}

int lbm_preset_poiseuille(lbm_domain* d, double rho_in, double rho_out)
{
  Rules r(d);
  pressure_rows(r, rho_in, rho_out);
  bounce_back_columns(r);
  return r.commit();
}

int lbm_preset_specular_channel(lbm_domain* d, double rho_in, double rho_out)
{
  Rules r(d);
  pressure_rows(r, rho_in, rho_out);
  specular_columns(r);
  return r.commit();
}

int lbm_preset_free_stream(lbm_domain* d, double uwx, double uwy)
{
  Rules r(d);
  // inlet row 0 and outlet row -1: all eight moving populations, fixed u_w
  for (int row = 0; row >= -1; row--)
  {
    lbm_bc_op op = Rules::base(0, row, one_end(row), 0, LBM_END);
    op.kind = LBM_BC_ABB_FIXED; op.src_q = -1; op.uw[0] = uwx; op.uw[1] = uwy;
    r.add(op);
  }
  specular_columns(r);
  return r.commit();
}

int lbm_preset_sedimentation(lbm_domain* d, double u_lb, const double* C_w, int R23, int C28, int C38)
{
  Rules r(d);
  // --- pre-stream, sediment lattice: zero gradient (top row from row 1, then outlet column from column -2)
  lbm_bc_op op = Rules::base(1, 0, 1, 0, LBM_END);
  op.kind = LBM_BC_COPY_PRE; op.src_mode = LBM_SRC_ROW; op.src_a = 1;
  r.add(op);
  op = Rules::base(1, 1, -1, -1, LBM_END);
  op.kind = LBM_BC_COPY_PRE; op.src_mode = LBM_SRC_COL; op.src_a = -2;
  r.add(op);
  // --- post-stream, fluid lattice
  op = Rules::base(0, 1, -1, 0, 1);  // inlet: rows 1..-2 of column 0, u_w = (0, u_lb)
  op.kind = LBM_BC_ABB_FIXED; op.src_q = -1; op.uw[0] = 0.0; op.uw[1] = u_lb;
  r.add(op);
  op = Rules::base(0, 0, LBM_END, -1, LBM_END);  // outlet: every row of the last column, extrapolated u_w
  op.kind = LBM_BC_ABB_EXTRAPOLATED; op.src_q = -1;
  r.add(op);
  // specular top
  r.local(0, 0, 1, 0, LBM_END, 8, 7);
  r.local(0, 0, 1, 0, LBM_END, 1, 3);
  r.local(0, 0, 1, 0, LBM_END, 5, 6);
  // no-slip bottom
  r.local(0, -1, LBM_END, 0, LBM_END, 7, 5);
  r.local(0, -1, LBM_END, 0, LBM_END, 3, 1);
  r.local(0, -1, LBM_END, 0, LBM_END, 6, 8);
  // rectangle: first wall, ceiling, second wall
  r.local(0, R23 + 1, -1, C28, C28 + 1, 8, 6);
  r.local(0, R23 + 1, -1, C28, C28 + 1, 4, 2);
  r.local(0, R23 + 1, -1, C28, C28 + 1, 7, 5);
  r.local(0, R23, R23 + 1, C28, C38 + 1, 6, 8);
  r.local(0, R23, R23 + 1, C28, C38 + 1, 3, 1);
  r.local(0, R23, R23 + 1, C28, C38 + 1, 7, 5);
  r.local(0, R23 + 1, -1, C38, C38 + 1, 5, 7);
  r.local(0, R23 + 1, -1, C38, C38 + 1, 2, 4);
  r.local(0, R23 + 1, -1, C38, C38 + 1, 6, 8);
  // --- post-stream, sediment lattice
  op = Rules::base(1, 1, -1, 0, 1);
  op.kind = LBM_BC_ADE_INLET; op.src_q = -1; op.per_row = C_w;
  r.add(op);
  r.local(1, R23 + 1, LBM_END, C28, C28 + 1, 8, 6, -1.0);
  r.local(1, R23 + 1, LBM_END, C28, C28 + 1, 4, 2, -1.0);
  r.local(1, R23 + 1, LBM_END, C28, C28 + 1, 7, 5, -1.0);
  r.local(1, R23, R23 + 1, C28, C38 + 1, 6, 8, -1.0);
  r.local(1, R23, R23 + 1, C28, C38 + 1, 3, 1, -1.0);
  r.local(1, R23, R23 + 1, C28, C38 + 1, 7, 5, -1.0);
  r.local(1, R23 + 1, -1, C38, C38 + 1, 5, 7, -1.0);
  r.local(1, R23 + 1, -1, C38, C38 + 1, 2, 4, -1.0);
  r.local(1, R23 + 1, -1, C38, C38 + 1, 6, 8, -1.0);
  r.local(1, -1, LBM_END, 0, LBM_END, 6, 8);
  r.local(1, -1, LBM_END, 0, LBM_END, 3, 1);
  r.local(1, -1, LBM_END, 0, LBM_END, 7, 5);
  return r.commit();
}

int lbm_preset_mrtcg(lbm_domain* d)
{
  Rules r(d);
  // "inlet-outlet": rows 1..-2, same-row copy from the opposite column (no diagonal shift)
  r.from(-1, 1, -1, 0, 1, 2, LBM_SRC_COL, -1);
  r.from(-1, 1, -1, 0, 1, 5, LBM_SRC_COL, -1);
  r.from(-1, 1, -1, 0, 1, 6, LBM_SRC_COL, -1);
  r.from(-1, 1, -1, -1, LBM_END, 4, LBM_SRC_COL, 0);
  r.from(-1, 1, -1, -1, LBM_END, 8, LBM_SRC_COL, 0);
  r.from(-1, 1, -1, -1, LBM_END, 7, LBM_SRC_COL, 0);
  // half-way bounce-back on the last and first row
  r.local(-1, -1, LBM_END, 0, LBM_END, 3, 1);
  r.local(-1, -1, LBM_END, 0, LBM_END, 7, 5);
  r.local(-1, -1, LBM_END, 0, LBM_END, 6, 8);
  r.local(-1, 0, 1, 0, LBM_END, 1, 3);
  r.local(-1, 0, 1, 0, LBM_END, 5, 7);
  r.local(-1, 0, 1, 0, LBM_END, 8, 6);
  return r.commit();
}

int lbm_preset_rk(lbm_domain* d)
{
  Rules r(d);
  // adv[left] = col[right]; adv[right] = col[left]; adv[top] = col[bottom]; adv[bottom] = col[top]
  r.from(-1, 1, -1, 0, 1, -1, LBM_SRC_COL, -1);
  r.from(-1, 1, -1, -1, LBM_END, -1, LBM_SRC_COL, 0);
  r.from(-1, 0, 1, 0, LBM_END, -1, LBM_SRC_ROW, -1);
  r.from(-1, -1, LBM_END, 0, LBM_END, -1, LBM_SRC_ROW, 0);
  return r.commit();
}

}  // extern "C"
