// lbm_presets.cpp — the boundary rule lists of the reference's drivers, written against the public
// C ABI only (lbm_bc_add), one lbm_bc_op per slice assignment and in the drivers' order so that
// later assignments win at corners exactly as they do there.
#include "../../include/lbm_b200.h"

namespace lbm
{
void set_error(const char* fmt, ...);
}

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

// Half-way bounce-back on every lattice link that joins a solid and a non-solid node: population q arriving at node n
// from n - c_q on the other side of the surface is replaced by the node's own outgoing opposite population,
//   f_adve[n, q] = f_coll[n, opp(q)]
// — the rule the reference writes out slice by slice for its axis-aligned obstacle walls
// (test/rectangle_sedimentation_test.cpp:186-196), applied to an arbitrary (staircase) body.  Both sides of the
// surface get the rule, so the solid region is a closed enclosure that exchanges nothing with the flow.
// Neighbours wrap periodically like solver::advect.  One LBM_BC_LINEAR op per run of consecutive columns.
int lbm_bc_add_solid(lbm_domain* d, int lattice, const unsigned char* solid, int X, int Y)
{
  if (!d || !solid || X < 1 || Y < 1) { lbm::set_error("lbm_bc_add_solid: null argument or empty mask"); return LBM_ERR_INVALID; }
  static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
  static const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
  for (int x = 0; x < X; x++)
    for (int q = 1; q < 9; q++)
    {
      const int sx = (x - cx[q] + X) % X;
      int run = -1;
      for (int y = 0; y <= Y; y++)
      {
        const bool cut = y < Y && (solid[(size_t)x * Y + y] != 0) != (solid[(size_t)sx * Y + (y - cy[q] + Y) % Y] != 0);
        if (cut && run < 0) run = y;
        if (!cut && run >= 0)
        {
          lbm_bc_op op;
          lbm_bc_op_default(&op);
          op.kind = LBM_BC_LINEAR; op.lattice = lattice;
          op.x_begin = x; op.x_end = x + 1; op.y_begin = run; op.y_end = y;
          op.dst_q = q; op.src_q = opp[q]; op.coef = 1.0;
          const int s = lbm_bc_add(d, &op);
          if (s != LBM_OK) return s;
          run = -1;
        }
      }
    }
  return LBM_OK;
}

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
