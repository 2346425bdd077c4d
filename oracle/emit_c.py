"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

C emitter restating the *output format* of pyoomph's code generator for the assembly routines:
  FiniteElementCode::write_code            /root/reference/src/codegen.cpp:4684-4811 (file skeleton)
  write_generic_RJM                        src/codegen.cpp:3912-4109       (loop nest)
  write_nodal_time_interpolation           src/codegen.cpp:1202-1292
  write_spatial_interpolation (+COORDDIFF) src/codegen.cpp:1294-1448
  write_generic_RJM_contribution           src/codegen.cpp:2015-2236
  write_generic_RJM_jacobian_contribution  src/codegen.cpp:1883-2013
  write_code_info (JIT_ELEMENT_init)       src/codegen.cpp:6398-7090

The symbolic residual comes from the shared front end (pyoomph_b200.codegen.FiniteElementCode, the stand-in
for pyoomph's GiNaC trees).  The Jacobian is derived HERE, independently of the product's coefficient-form
derivation: like the reference it is the derivative of the complete residual expression with respect to the
nodal value U^{l_shape} of one field (a Gateaux derivative through every shape expansion, the test-function
gradients and the integration weight), printed as one expression per (test field, unknown field) and evaluated
for every (l_test, l_shape) pair.  Moving-mesh terms use the reference's tensors
(``d_dx_shape_dcoord_*``, ``int_pt_weights_d_coords``, COORDDIFF arrays), not the product's closed-form identity.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Tuple

import sympy as sp
from sympy.printing.c import C99CodePrinter

from pyoomph_b200 import expressions as ex
from pyoomph_b200.codegen import AtomInfo, FiniteElementCode

EPS = sp.Symbol("ORACLE__eps", real=True)
MM = sp.Symbol("ORACLE__partial_t_mass_matrix", real=True)


class _CPrinter(C99CodePrinter):
    def __init__(self, names: Dict[sp.Symbol, str]):
        super().__init__({"precision": 17})
        self.names = names

    def _print_Symbol(self, expr):
        if expr in self.names:
            return self.names[expr]
        return super()._print_Symbol(expr)

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Rational(self, expr):
        return "(%d.0/%d.0)" % (expr.p, expr.q)

    def _print_Integer(self, expr):
        return "%d.0" % int(expr) if abs(int(expr)) < 2 ** 53 else super()._print_Integer(expr)

    def _print_Pow(self, expr):
        # integer exponents stay integers (the base rule above would print them as doubles)
        if expr.exp.is_Integer:
            e = int(expr.exp)
            b = self.parenthesize(expr.base, sp.printing.precedence.PRECEDENCE["Pow"])
            if e == -1:
                return "1.0/%s" % b
            return "pow(%s,%d.0)" % (self._print(expr.base), e)
        return super()._print_Pow(expr)


def _space_shape_name(code: FiniteElementCode, fieldname: str) -> str:
    return {"C2": "C2", "C1": "C1", "Pos": "Pos"}[code.fields[fieldname].space]


def _nodal_index_name(fieldname: str) -> str:
    return "this_nodalind_" + fieldname


class CEmitter:
    def __init__(self, code: FiniteElementCode):
        self.code = code
        self.dim = code.nodal_dim

    # ---- naming helpers ---------------------------------------------------------
    def _data_array(self, field: str) -> str:
        return "nodal_coords" if self.code.fields[field].space == "Pos" else "nodal_data"

    def _eqn_str(self, field: str, l: str) -> str:
        arr = "pos_local_eqn" if self.code.fields[field].space == "Pos" else "nodal_local_eqn"
        return "eleminfo->%s[%s][%s]" % (arr, l, _nodal_index_name(field))

    def _nnode_str(self, field: str) -> str:
        s = self.code.fields[field].space
        return "eleminfo->nnode" if s == "Pos" else "eleminfo->nnode_" + s

    def _shape_str(self, field: str, deriv: str, l: str) -> str:
        S = _space_shape_name(self.code, field)
        if deriv == "d0":
            return "shapeinfo->shape_%s[%s]" % (S, l)
        return "shapeinfo->d%s_shape_%s[%s][%s]" % (deriv[1], S, l, deriv[2:])

    def _dt_values_name(self, a: AtomInfo) -> str:
        return "this_d%dt%d%s_%s" % (a.dt_order, a.past, a.scheme.replace("_degr", ""), a.field)

    def _weights_name(self, a: AtomInfo) -> str:
        return "shapeinfo->timestepper_weights_%s_%s" % ("dt" if a.dt_order == 1 else "d2t", a.scheme)

    def _coorddiff_name(self, a: AtomInfo, j: int) -> str:
        dts = "d%dt%d%s" % (a.dt_order, a.past, a.scheme.replace("_degr", "") if a.dt_order else "")
        return "this_intrp_%s_d1x%s_COORDDIFF_%d_%s" % (dts, a.deriv[2:], j, a.field)

    # ---- one routine ---------------------------------------------------------------------------
    def routine(self, funcname: str, resname: str, res_index: int, parameter: Optional[str]) -> str:
        code = self.code
        E = code.atomize(code.residuals[resname])
        if parameter is not None:
            E = sp.diff(E, code._param_syms[parameter])
        atoms = sorted([code._atom_syms[s] for s in E.free_symbols if s in code._atom_syms],
                       key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        moving = code.coordinates_as_dofs
        if moving:
            # Eulerian-gradient atoms need all directions for nothing extra here (tensors do the work)
            pass
        tests = sorted([code._test_syms[s] for s in E.free_symbols if s in code._test_syms], key=lambda t: (t.field, t.deriv))
        test_fields = []
        for t in tests:
            if t.field not in test_fields:
                test_fields.append(t.field)
        names: Dict[sp.Symbol, str] = {ex.DX_EUL: "dx", ex.DX_LAG: "dX", ex.TIME: "t[0]", ex.pi: "Pi",
                                       ex.NORMAL[0]: "shapeinfo->normal[0]", ex.NORMAL[1]: "shapeinfo->normal[1]", ex.NORMAL[2]: "shapeinfo->normal[2]",
                                       ex.ELEMSIZE_EUL: "shapeinfo->elemsize_Eulerian", ex.ELEMSIZE_EUL_CART: "shapeinfo->elemsize_Eulerian_cartesian",
                                       ex.ELEMSIZE_LAG: "shapeinfo->elemsize_Lagrangian", ex.ELEMSIZE_LAG_CART: "shapeinfo->elemsize_Lagrangian_cartesian"}
        for a in atoms:
            names[code.atom_symbol(a)] = "this_" + a.cname
        for k, p in enumerate(code.global_params):
            names[code._param_syms[p]] = "(*(my_func_table->global_parameters[%d]))" % k
        for s in code._test_syms:
            sl = code._test_syms[s]
            names[s] = "testfunction[l_test]" if sl.deriv == "d0" else "d%s_testfunction[l_test][%s]" % (sl.deriv[1], sl.deriv[2:])
        pr = _CPrinter(names)

        o: List[str] = []
        w = o.append
        w("static void %s(const JITElementInfo_t * eleminfo, const JITShapeInfo_t * shapeinfo,double * residuals, double *jacobian, double *mass_matrix,unsigned flag)" % funcname)
        w("{")
        w("  int local_eqn, local_unknown;")
        w("  unsigned nummaster,nummaster2;")
        w("  double hang_weight,hang_weight2;")
        w("  const double * t=shapeinfo->t;")
        w("  const double * dt=shapeinfo->dt;")
        w("  (void)t; (void)dt; (void)local_unknown; (void)nummaster2; (void)hang_weight2;")
        idx_fields = sorted({a.field for a in atoms} | set(test_fields) |
                            ({"coordinate_" + d for d in ex.DIRS[:self.dim]} if moving else set()))
        for f in idx_fields:
            w("  const unsigned %s = %d;" % (_nodal_index_name(f), code.fields[f].index))
        w("  //START: Precalculate time derivatives of the necessary data")
        dt_atoms: Dict[str, AtomInfo] = {}
        for a in atoms:
            if a.dt_order:
                dt_atoms.setdefault(self._dt_values_name(a), a)
        by_space: Dict[str, List[Tuple[str, AtomInfo]]] = {}
        for nm, a in dt_atoms.items():
            by_space.setdefault(code.fields[a.field].space, []).append((nm, a))
        for space, lst in by_space.items():
            rng = self._nnode_str(lst[0][1].field)
            for nm, a in lst:
                w("  double %s[%d];" % (nm, code.etype.nnode))
            w("  for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % rng)
            w("  {")
            for nm, a in lst:
                w("    %s[l_shape]=0.0;" % nm)
            w("    for (unsigned tindex=0;tindex<shapeinfo->timestepper_ntstorage;tindex++)")
            w("    {")
            for nm, a in lst:
                w("      %s[l_shape] += %s[tindex]*eleminfo->%s[l_shape][%s][tindex];" % (
                    nm, self._weights_name(a), self._data_array(a.field), _nodal_index_name(a.field)))
            w("    }")
            w("  }")
        w("  //END: Precalculate time derivatives of the necessary data")
        w("")
        w("  //START: Spatial integration loop")
        w("  for(unsigned ipt=0;ipt<shapeinfo->n_int_pt;ipt++)")
        w("  {")
        w("    my_func_table->fill_shape_buffer_for_point(ipt, &(my_func_table->shapes_required_ResJac[%d]), flag);" % res_index)
        w("    const double dx = shapeinfo->int_pt_weight;")
        w("    const double dX = shapeinfo->int_pt_weight_Lagrangian;")
        w("    (void)dx; (void)dX;")
        w("    //START: Interpolate all required fields")
        for space in ("Pos", "C2", "C1"):
            sat = [a for a in atoms if code.fields[a.field].space == space]
            if not sat:
                continue
            rng = self._nnode_str(sat[0].field)
            for a in sat:
                w("    double this_%s=0.0;" % a.cname)
            w("    for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % rng)
            w("    {")
            for a in sat:
                if a.dt_order:
                    nd = "%s[l_shape]" % self._dt_values_name(a)
                else:
                    nd = "eleminfo->%s[l_shape][%s][%d]" % (self._data_array(a.field), _nodal_index_name(a.field), a.past)
                w("      this_%s+= %s * %s;" % (a.cname, nd, self._shape_str(a.field, a.deriv, "l_shape")))
            w("    }")
            cd = [a for a in sat if moving and a.deriv.startswith("dx") and space != "Pos" and not a.past]
            if cd:
                for a in cd:
                    for j in range(self.dim):
                        w("    double %s[%d];" % (self._coorddiff_name(a, j), code.etype.nnode))
                w("    if (flag)")
                w("    {")
                w("     for (unsigned int m=0;m<eleminfo->nnode;m++)")
                w("     {")
                for a in cd:
                    for j in range(self.dim):
                        w("        %s[m]=0.0;" % self._coorddiff_name(a, j))
                w("        for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % rng)
                w("        {")
                for a in cd:
                    nd = ("%s[l_shape]" % self._dt_values_name(a)) if a.dt_order else \
                        "eleminfo->nodal_data[l_shape][%s][0]" % _nodal_index_name(a.field)
                    S = _space_shape_name(code, a.field)
                    for j in range(self.dim):
                        w("           %s[m]+=%s * shapeinfo->d_dx_shape_dcoord_%s[l_shape][%s][m][%d];" % (
                            self._coorddiff_name(a, j), nd, S, a.deriv[2:], j))
                w("        }")
                w("     }")
                w("    }")
        w("    //END: Interpolate all required fields")
        w("")
        w("    //START: Contribution of the spaces")
        w("    double _res_contrib,_J_contrib;")
        for space in ("Pos", "C2", "C1"):
            tf = [f for f in test_fields if code.fields[f].space == space]
            if not tf:
                continue
            S = {"Pos": "Pos"}.get(space, space)
            nn = "eleminfo->nnode" if space == "Pos" else "eleminfo->nnode_" + space
            w("    {")
            w("      double const * testfunction = shapeinfo->shape_%s;" % S)
            w("      DX_SHAPE_FUNCTION_DECL(dx_testfunction) = shapeinfo->dx_shape_%s;" % S)
            w("      DX_SHAPE_FUNCTION_DECL(dX_testfunction) = shapeinfo->dX_shape_%s;" % S)
            w("      (void)testfunction; (void)dx_testfunction; (void)dX_testfunction;")
            w("      for (unsigned int l_test=0;l_test<%s;l_test++)" % nn)
            w("      {")
            for F in tf:
                other = {s: 0 for s, sl in code._test_syms.items() if sl.field != F}
                var_part = E.xreplace(other)
                if var_part == 0:
                    continue
                hang = "shapeinfo->hanginfo_%s" % S
                w("        BEGIN_RESIDUAL_CONTINUOUS_SPACE(%s,%s, %s,%s,l_test)" % (
                    self._eqn_str(F, "l_test"), pr.doprint(var_part), hang, _nodal_index_name(F)))
                w("          ADD_TO_RESIDUAL_CONTINUOUS_SPACE()")
                w("          BEGIN_JACOBIAN()")
                for G in code.unknown_field_names():
                    diffpart = self._gateaux(var_part, F, G, names)
                    if diffpart == 0:
                        continue
                    mass_part = sp.diff(diffpart, MM)
                    diffpart = diffpart.xreplace({MM: 0})
                    Gs = _space_shape_name(code, G)
                    w("            for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % self._nnode_str(G))
                    w("            {")
                    w("              BEGIN_JACOBIAN_HANG(%s, %s,shapeinfo->hanginfo_%s,%s,l_shape)" % (
                        self._eqn_str(G, "l_shape"), pr.doprint(diffpart), Gs, _nodal_index_name(G)))
                    w("                ADD_TO_JACOBIAN_HANG_HANG()")
                    if mass_part != 0:
                        w("                ADD_TO_MASS_MATRIX_HANG_HANG(%s)" % pr.doprint(mass_part))
                    w("              END_JACOBIAN_HANG()")
                    w("            }")
                w("          END_JACOBIAN()")
                w("        END_RESIDUAL_CONTINUOUS_SPACE()")
            w("      }")
            w("    }")
        w("    //END: Contribution of the spaces")
        w("  }")
        w("  //END: Spatial integration loop")
        w("}")
        w("")
        return "\n".join(o)

    # ---- integral expressions ---------------------------------------------------------------------
    def integral_routine(self) -> str:
        """EvalIntegralExpression in the format of write_code_integral_or_local_expressions (src/codegen.cpp:4125-4364,
        integrate = true)"""
        code = self.code
        return self._expression_routine("EvalIntegralExpression", list(code.integral_expressions.keys()),
                                        list(code.integral_expressions.values()), "IntegralExprs", integrate=True)

    def point_routines(self) -> str:
        """EvalLocalExpression / EvalExtremumExpression (the same writer with integrate = false: the host has filled the shape buffer
        at the local coordinate it wants, src/codegen.cpp:4181-4186, src/elements.cpp:4666-4704) and GetZ2Fluxes
        (write_code_get_z2_flux, src/codegen.cpp:4445-4530: all flux terms of the point into Z2Flux[])"""
        code = self.code
        out = []
        if code.local_expressions:
            out.append(self._expression_routine("EvalLocalExpression", list(code.local_expressions.keys()), list(code.local_expressions.values()),
                                                "LocalExprs", integrate=False))
        if code.extremum_expressions:
            out.append(self._expression_routine("EvalExtremumExpression", list(code.extremum_expressions.keys()),
                                                list(code.extremum_expressions.values()), "ExtremumExprs", integrate=False))
        if code.Z2_fluxes:
            out.append(self._expression_routine("GetZ2Fluxes", ["flux_%d" % i for i in range(len(code.Z2_fluxes))], list(code.Z2_fluxes),
                                                "Z2Fluxes", integrate=False, z2=True))
        return "\n".join(out)

    def _expression_routine(self, funcname: str, enames: List[str], raw_exprs, reqname: str, integrate: bool, z2: bool = False) -> str:
        code = self.code
        exprs = [code.atomize(e) for e in raw_exprs]
        used = set()
        for e in exprs:
            used |= {s_ for s_ in e.free_symbols if s_ in code._atom_syms}
        atoms = sorted([code._atom_syms[s_] for s_ in used], key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        names: Dict[sp.Symbol, str] = {ex.DX_EUL: "dx", ex.DX_LAG: "dX", ex.TIME: "t[0]", ex.pi: "Pi",
                                       ex.NORMAL[0]: "shapeinfo->normal[0]", ex.NORMAL[1]: "shapeinfo->normal[1]", ex.NORMAL[2]: "shapeinfo->normal[2]",
                                       ex.ELEMSIZE_EUL: "shapeinfo->elemsize_Eulerian", ex.ELEMSIZE_EUL_CART: "shapeinfo->elemsize_Eulerian_cartesian",
                                       ex.ELEMSIZE_LAG: "shapeinfo->elemsize_Lagrangian", ex.ELEMSIZE_LAG_CART: "shapeinfo->elemsize_Lagrangian_cartesian"}
        for a in atoms:
            names[code.atom_symbol(a)] = "this_" + a.cname
        for k, p in enumerate(code.global_params):
            names[code._param_syms[p]] = "(*(my_func_table->global_parameters[%d]))" % k
        pr = _CPrinter(names)
        o: List[str] = []
        w = o.append
        if z2:
            w("static void %s(const JITElementInfo_t * eleminfo, const JITShapeInfo_t * shapeinfo, double * Z2Flux)" % funcname)
        else:
            w("static double %s(const JITElementInfo_t * eleminfo, const JITShapeInfo_t * shapeinfo, unsigned index)" % funcname)
        w("{")
        w("  const unsigned flag=0;")
        w("  const double * t=shapeinfo->t;")
        w("  const double * dt=shapeinfo->dt;")
        w("  (void)t; (void)dt; (void)flag;")
        for f in sorted({a.field for a in atoms}):
            w("  const unsigned %s = %d;" % (_nodal_index_name(f), code.fields[f].index))
        w("  //START: Precalculate time derivatives of the necessary data")
        dt_atoms: Dict[str, AtomInfo] = {}
        for a in atoms:
            if a.dt_order:
                dt_atoms.setdefault(self._dt_values_name(a), a)
        by_space: Dict[str, List[Tuple[str, AtomInfo]]] = {}
        for nm, a in dt_atoms.items():
            by_space.setdefault(code.fields[a.field].space, []).append((nm, a))
        for space, lst in by_space.items():
            rng = self._nnode_str(lst[0][1].field)
            for nm, a in lst:
                w("  double %s[%d];" % (nm, code.etype.nnode))
            w("  for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % rng)
            w("  {")
            for nm, a in lst:
                w("    %s[l_shape]=0.0;" % nm)
            w("    for (unsigned tindex=0;tindex<shapeinfo->timestepper_ntstorage;tindex++)")
            w("    {")
            for nm, a in lst:
                w("      %s[l_shape] += %s[tindex]*eleminfo->%s[l_shape][%s][tindex];" % (
                    nm, self._weights_name(a), self._data_array(a.field), _nodal_index_name(a.field)))
            w("    }")
            w("  }")
        w("  //END: Precalculate time derivatives of the necessary data")
        w("")
        if integrate:
            w("  double res=0.0;")
            w("  for(unsigned ipt=0;ipt<shapeinfo->n_int_pt;ipt++)")
            w("  {")
            w("    my_func_table->fill_shape_buffer_for_point(ipt, &(my_func_table->shapes_required_%s), 0);" % reqname)
        else:
            if not z2:
                w("  double res;")
            w("  unsigned ipt=0; (void)ipt;")
            w("  {")
        w("    const double dx = shapeinfo->int_pt_weight;")
        w("    const double dX = shapeinfo->int_pt_weight_Lagrangian;")
        w("    (void)dx; (void)dX;")
        w("    //START: Interpolate all required fields")
        for space in ("Pos", "C2", "C1"):
            sat = [a for a in atoms if code.fields[a.field].space == space]
            if not sat:
                continue
            rng = self._nnode_str(sat[0].field)
            for a in sat:
                w("    double this_%s=0.0;" % a.cname)
            w("    for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % rng)
            w("    {")
            for a in sat:
                if a.dt_order:
                    nd = "%s[l_shape]" % self._dt_values_name(a)
                else:
                    nd = "eleminfo->%s[l_shape][%s][%d]" % (self._data_array(a.field), _nodal_index_name(a.field), a.past)
                w("      this_%s+= %s * %s;" % (a.cname, nd, self._shape_str(a.field, a.deriv, "l_shape")))
            w("    }")
        w("    //END: Interpolate all required fields")
        if z2:
            for i, (n, e) in enumerate(zip(enames, exprs)):
                w("    Z2Flux[%d] = %s; // %s" % (i, pr.doprint(e), n))
        else:
            w("    switch (index)")
            w("    {")
            for i, (n, e) in enumerate(zip(enames, exprs)):
                w("      case %d: res%s= %s; break; // %s" % (i, "+" if integrate else "", pr.doprint(e), n))
            if not integrate:
                w("      default: res=0.0;")
            w("    }")
        w("  }")
        if not z2:
            w("  return res;")
        w("}")
        w("")
        return "\n".join(o)

    def _gateaux(self, var_part: sp.Expr, F: str, G: str, names: Dict[sp.Symbol, str], lname: str = "l_shape", mass: bool = True) -> sp.Expr:
        """d/d U_G^{lname} of the complete residual expression of test field F."""
        code = self.code
        var: Dict[sp.Symbol, sp.Expr] = {}

        def csym(text: str) -> sp.Symbol:
            s = sp.Symbol("C__" + text, real=True)
            names[s] = text
            return s

        for s in list(var_part.free_symbols):
            if s in code._atom_syms:
                a = code._atom_syms[s]
                if a.past:
                    continue
                if a.field == G:
                    shp = csym(self._shape_str(G, a.deriv, lname))
                    if a.dt_order == 0:
                        var[s] = shp
                    else:
                        var[s] = (csym(self._weights_name(a) + "[0]") + (MM if (a.dt_order == 1 and mass) else 0)) * shp
                elif G.startswith("coordinate_") and a.deriv.startswith("dx") and code.fields[a.field].space != "Pos":
                    j = ex.DIRS.index(G[-1])
                    var[s] = csym(self._coorddiff_name(a, j) + "[%s]" % lname)
        if G.startswith("coordinate_") and code.coordinates_as_dofs:
            j = ex.DIRS.index(G[-1])
            if var_part.has(ex.DX_EUL):
                var[ex.DX_EUL] = csym("shapeinfo->int_pt_weights_d_coords[%d][l_shape]" % j)
            for i in range(self.dim):     # interface elements: the normal moves with the nodes (GiNaCNormalSymbol derivative -> d_normal_dcoord)
                if var_part.has(ex.NORMAL[i]):
                    var[ex.NORMAL[i]] = csym("shapeinfo->d_normal_dcoord[%d][l_shape][%d]" % (i, j))
            for s, sl in code._test_syms.items():
                if sl.field == F and sl.deriv.startswith("dx") and var_part.has(s):
                    S = _space_shape_name(code, F)
                    var[s] = csym("shapeinfo->d_dx_shape_dcoord_%s[l_test][%s][l_shape][%d]" % (S, sl.deriv[2:], j))
        if not var:
            return sp.Integer(0)
        perturbed = var_part.xreplace({s: s + EPS * v for s, v in var.items()})
        return sp.diff(perturbed, EPS).xreplace({EPS: 0})

    # ---- HessianVectorProduct<i> (src/codegen.cpp:3646-3910, :1500-1881) ------------------------
    def hessian_routine(self, funcname: str, resname: str, res_index: int) -> str:
        """Full (non-symmetric) assembly of H[i][j][k] = d/dU_k dR_i/dU_j and of the mass Hessian into n_dof^3 buffers,
        followed by the reference's tail contractions (src/codegen.cpp:3879-3905, macros src/jitbridge.h:624-691)."""
        code = self.code
        if code.coordinates_as_dofs:
            raise RuntimeError("oracle Hessian: fixed meshes only")
        E = code.atomize(code.residuals[resname])
        atoms = sorted([code._atom_syms[s] for s in E.free_symbols if s in code._atom_syms], key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        tests = sorted([code._test_syms[s] for s in E.free_symbols if s in code._test_syms], key=lambda t: (t.field, t.deriv))
        test_fields = []
        for t in tests:
            if t.field not in test_fields:
                test_fields.append(t.field)
        names: Dict[sp.Symbol, str] = {ex.DX_EUL: "dx", ex.DX_LAG: "dX", ex.TIME: "t[0]", ex.pi: "Pi",
                                       ex.NORMAL[0]: "shapeinfo->normal[0]", ex.NORMAL[1]: "shapeinfo->normal[1]", ex.NORMAL[2]: "shapeinfo->normal[2]",
                                       ex.ELEMSIZE_EUL: "shapeinfo->elemsize_Eulerian", ex.ELEMSIZE_EUL_CART: "shapeinfo->elemsize_Eulerian_cartesian",
                                       ex.ELEMSIZE_LAG: "shapeinfo->elemsize_Lagrangian", ex.ELEMSIZE_LAG_CART: "shapeinfo->elemsize_Lagrangian_cartesian"}
        for a in atoms:
            names[code.atom_symbol(a)] = "this_" + a.cname
        for k, p in enumerate(code.global_params):
            names[code._param_syms[p]] = "(*(my_func_table->global_parameters[%d]))" % k
        for s in code._test_syms:
            sl = code._test_syms[s]
            names[s] = "testfunction[l_test]" if sl.deriv == "d0" else "d%s_testfunction[l_test][%s]" % (sl.deriv[1], sl.deriv[2:])
        pr = _CPrinter(names)
        o: List[str] = []
        w = o.append
        w("static void %s(const JITElementInfo_t * eleminfo, const JITShapeInfo_t * shapeinfo,const double * Y, double * Cs, double * product, unsigned numvectors, unsigned flag)" % funcname)
        w("{")
        w("  int local_eqn, local_unknown, local_deriv;")
        w("  const double * t=shapeinfo->t;")
        w("  const double * dt=shapeinfo->dt;")
        w("  (void)t; (void)dt;")
        w("  const unsigned n_dof=shapeinfo->jacobian_size;")
        w("  double * hessian_buffer=(flag==3 ? product : (double*)calloc(n_dof*n_dof*n_dof,sizeof(double)));")
        w("  double * hessian_M_buffer=(flag==3 ? Cs : ((flag==2 || flag==5) ? (double*)calloc(n_dof*n_dof*n_dof,sizeof(double)) : PYOOMPH_NULL_));")
        for f in sorted({a.field for a in atoms} | set(test_fields)):
            w("  const unsigned %s = %d;" % (_nodal_index_name(f), code.fields[f].index))
        dt_atoms: Dict[str, AtomInfo] = {}
        for a in atoms:
            if a.dt_order:
                dt_atoms.setdefault(self._dt_values_name(a), a)
        for nm, a in dt_atoms.items():
            w("  double %s[%d];" % (nm, code.etype.nnode))
            w("  for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % self._nnode_str(a.field))
            w("  {")
            w("    %s[l_shape]=0.0;" % nm)
            w("    for (unsigned tindex=0;tindex<shapeinfo->timestepper_ntstorage;tindex++) %s[l_shape] += %s[tindex]*eleminfo->%s[l_shape][%s][tindex];" % (
                nm, self._weights_name(a), self._data_array(a.field), _nodal_index_name(a.field)))
            w("  }")
        w("  for(unsigned ipt=0;ipt<shapeinfo->n_int_pt;ipt++)")
        w("  {")
        w("    my_func_table->fill_shape_buffer_for_point(ipt, &(my_func_table->shapes_required_Hessian[%d]), 3);" % res_index)
        w("    const double dx = shapeinfo->int_pt_weight;")
        w("    const double dX = shapeinfo->int_pt_weight_Lagrangian;")
        w("    (void)dx; (void)dX;")
        for space in ("Pos", "C2", "C1"):
            sat = [a for a in atoms if code.fields[a.field].space == space]
            if not sat:
                continue
            for a in sat:
                w("    double this_%s=0.0;" % a.cname)
            w("    for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % self._nnode_str(sat[0].field))
            w("    {")
            for a in sat:
                nd = ("%s[l_shape]" % self._dt_values_name(a)) if a.dt_order else "eleminfo->%s[l_shape][%s][%d]" % (self._data_array(a.field), _nodal_index_name(a.field), a.past)
                w("      this_%s+= %s * %s;" % (a.cname, nd, self._shape_str(a.field, a.deriv, "l_shape")))
            w("    }")
        for space in ("C2", "C1"):
            tf = [f for f in test_fields if code.fields[f].space == space]
            if not tf:
                continue
            w("    {")
            w("      double const * testfunction = shapeinfo->shape_%s;" % space)
            w("      DX_SHAPE_FUNCTION_DECL(dx_testfunction) = shapeinfo->dx_shape_%s;" % space)
            w("      DX_SHAPE_FUNCTION_DECL(dX_testfunction) = shapeinfo->dX_shape_%s;" % space)
            w("      (void)testfunction; (void)dx_testfunction; (void)dX_testfunction;")
            w("      for (unsigned int l_test=0;l_test<eleminfo->nnode_%s;l_test++)" % space)
            w("      {")
            for F in tf:
                other = {s: 0 for s, sl in code._test_syms.items() if sl.field != F}
                var_part = E.xreplace(other)
                if var_part == 0:
                    continue
                w("        local_eqn=%s;" % self._eqn_str(F, "l_test"))
                w("        if (local_eqn>=0)")
                w("        {")
                for G in code.unknown_field_names():
                    d1 = self._gateaux(var_part, F, G, names, "l_shape")
                    if d1 == 0:
                        continue
                    m1 = sp.diff(d1, MM)
                    d1 = d1.xreplace({MM: 0})
                    for H in code.unknown_field_names():
                        d2 = self._gateaux(d1, F, H, names, "l_shape2", mass=False)
                        m2 = self._gateaux(m1, F, H, names, "l_shape2", mass=False) if m1 != 0 else sp.Integer(0)
                        if d2 == 0 and m2 == 0:
                            continue
                        w("          for (unsigned int l_shape=0;l_shape<%s;l_shape++)" % self._nnode_str(G))
                        w("          {")
                        w("            local_unknown=%s;" % self._eqn_str(G, "l_shape"))
                        w("            if (local_unknown>=0)")
                        w("            {")
                        w("              for (unsigned int l_shape2=0;l_shape2<%s;l_shape2++)" % self._nnode_str(H))
                        w("              {")
                        w("                local_deriv=%s;" % self._eqn_str(H, "l_shape2"))
                        w("                if (local_deriv>=0)")
                        w("                {")
                        if d2 != 0:
                            w("                  hessian_buffer[local_eqn*n_dof*n_dof+local_unknown*n_dof+local_deriv] += %s;" % pr.doprint(d2))
                        if m2 != 0:
                            w("                  if (flag>=2 && flag!=4) hessian_M_buffer[local_eqn*n_dof*n_dof+local_unknown*n_dof+local_deriv] += %s;" % pr.doprint(m2))
                        w("                }")
                        w("              }")
                        w("            }")
                        w("          }")
                w("        }")
            w("      }")
            w("    }")
        w("  }")
        # tail (jitbridge.h:637-691 restated: contraction of the middle index, or of the first one when transposed)
        w("  if (!flag)")
        w("  {")
        w("    for (unsigned int i=0;i<n_dof;i++) for (unsigned int k=0;k<n_dof;k++)")
        w("    {")
        w("      double Yj_Hijk=0.0;")
        w("      for (unsigned int j=0;j<n_dof;j++) Yj_Hijk+=Y[j]*hessian_buffer[i*n_dof*n_dof+j*n_dof+k];")
        w("      for (unsigned int v=0;v<numvectors;v++) product[v*n_dof+i]+=Yj_Hijk*Cs[v*n_dof+k];")
        w("    }")
        w("    free(hessian_buffer);")
        w("  }")
        w("  else if (flag!=3)")
        w("  {")
        w("    for (unsigned int ivec=0;ivec<numvectors;ivec++) for (unsigned i=0;i<n_dof;i++) for (unsigned k=0;k<n_dof;k++) for (unsigned int j=0;j<n_dof;j++)")
        w("    {")
        w("      if (flag==5 || flag==4) product[n_dof*n_dof*ivec+i*n_dof+k] += hessian_buffer[j*n_dof*n_dof+i*n_dof+k]*Y[n_dof*ivec+j];")
        w("      else product[n_dof*n_dof*ivec+i*n_dof+k] += hessian_buffer[i*n_dof*n_dof+j*n_dof+k]*Y[n_dof*ivec+j];")
        w("    }")
        w("    free(hessian_buffer);")
        w("  }")
        w("  if (flag==2 || flag==5)")
        w("  {")
        w("    for (unsigned int ivec=0;ivec<numvectors;ivec++) for (unsigned i=0;i<n_dof;i++) for (unsigned k=0;k<n_dof;k++) for (unsigned int j=0;j<n_dof;j++)")
        w("    {")
        w("      if (flag==5) Cs[n_dof*n_dof*ivec+i*n_dof+k] += hessian_M_buffer[j*n_dof*n_dof+i*n_dof+k]*Y[n_dof*ivec+j];")
        w("      else Cs[n_dof*n_dof*ivec+i*n_dof+k] += hessian_M_buffer[i*n_dof*n_dof+j*n_dof+k]*Y[n_dof*ivec+j];")
        w("    }")
        w("    free(hessian_M_buffer);")
        w("  }")
        w("}")
        w("")
        return "\n".join(o)

    # ---- whole plugin file -------------------------------------------------------------------
    def emit(self) -> str:
        code = self.code
        o: List[str] = []
        w = o.append
        w("/* generated by oracle/emit_c.py in the format of pyoomph's FiniteElementCode::write_code -- TEST INFRASTRUCTURE */")
        w("#define JIT_ELEMENT_SHARED_LIB")
        w('#include "oracle_jit.h"')
        w("static JITFuncSpec_Table_FiniteElement_t * my_func_table;")
        w('#include "oracle_jit_hang.h"')
        w("")
        resnames = code.residual_names()
        # src/codegen.cpp:5762-5772: the coordinate system's geometric Jacobian as a function of the position, used for elemsize_Eulerian
        w("// Used for elemsize_Eulerian etc")
        w("static double JacobianForElementSize(const JITElementInfo_t * eleminfo, const double * _x)")
        w("{")
        w("  return %s;" % ("2*Pi*_x[0]" if code.coordinate_system.get_id_name() == "Axisymmetric" else "1.0"))
        w("}")
        w("")
        for i, rn in enumerate(resnames):
            w(self.routine("ResidualAndJacobian%d" % i, rn, i, None))
            for p in code.global_params:
                w(self.routine("dResidual%ddParameter_%s" % (i, p), rn, i, p))
        hess = not code.coordinates_as_dofs
        if hess:
            w("#ifndef PYOOMPH_NULL_")
            w("#define PYOOMPH_NULL_ ((double*)0)")
            w("#endif")
            for i, rn in enumerate(resnames):
                w(self.hessian_routine("HessianVectorProduct%d" % i, rn, i))
        nint = len(code.integral_expressions)
        if code.point_expression_names():
            w(self.point_routines())
        if nint:
            w(self.integral_routine())
        nC2 = len([f for f in code.nodal_fields() if f.space == "C2"])
        nC1 = len([f for f in code.nodal_fields() if f.space == "C1"])
        w("static void clean_up(JITFuncSpec_Table_FiniteElement_t *functable)")
        w("{")
        if nint:
            w(" for (unsigned i=0;i<functable->numintegral_expressions;i++) { pyoomph_tested_free(functable->integral_expressions_names[i]); }")
            w(" pyoomph_tested_free(functable->integral_expressions_names);")
        w(" free(functable->ResidualAndJacobian); free(functable->ResidualAndJacobianSteady); free(functable->shapes_required_ResJac);")
        w(" if (functable->hessian_generated) { free(functable->HessianVectorProduct); free(functable->shapes_required_Hessian); }")
        w(" free(functable->global_parameters);")
        w(" for (unsigned i=0;i<functable->num_res_jacs;i++) { free(functable->ParameterDerivative[i]); free(functable->res_jac_names[i]); }")
        w(" free(functable->ParameterDerivative); free(functable->res_jac_names); free(functable->dominant_space);")
        w("}")
        w("")
        w("JIT_API void JIT_ELEMENT_init(JITFuncSpec_Table_FiniteElement_t *functable)")
        w("{")
        w(' functable->check_compiler_size(sizeof(double),8, "double");')
        w(" functable->nodal_dim=%d;" % self.dim)
        w(" functable->lagr_dim=%d;" % self.dim)
        w(" functable->fd_jacobian=false; ")
        w(" functable->fd_position_jacobian=false; ")
        w(" functable->with_adaptivity=true; ")
        w(" functable->numfields_C2=%d;" % nC2)
        w(" functable->numfields_C1=%d;" % nC1)
        w(" functable->numfields_Pos=%d;" % (2 * self.dim))
        w(" functable->num_res_jacs=%d;" % len(resnames))
        w(" functable->current_res_jac=0;")
        w(" functable->res_jac_names=(char **)calloc(%d,sizeof(char*));" % max(1, len(resnames)))
        for i, rn in enumerate(resnames):
            w(' SET_INTERNAL_NAME(functable->res_jac_names[%d],"%s");' % (i, rn))
        w(" functable->shapes_required_ResJac=(JITFuncSpec_RequiredShapes_FiniteElement_t *)calloc(%d,sizeof(JITFuncSpec_RequiredShapes_FiniteElement_t));" % max(1, len(resnames)))
        for i, rn in enumerate(resnames):
            if code.residuals[rn].has(ex.ELEMSIZE_EUL):
                w(" functable->shapes_required_ResJac[%d].elemsize_Eulerian_Pos=true;" % i)
            if code.residuals[rn].has(ex.ELEMSIZE_EUL_CART):
                w(" functable->shapes_required_ResJac[%d].elemsize_Eulerian_cartesian_Pos=true;" % i)
            if code.residuals[rn].has(ex.ELEMSIZE_LAG):
                w(" functable->shapes_required_ResJac[%d].elemsize_Lagrangian_Pos=true;" % i)
            if code.residuals[rn].has(ex.ELEMSIZE_LAG_CART):
                w(" functable->shapes_required_ResJac[%d].elemsize_Lagrangian_cartesian_Pos=true;" % i)
        w(" functable->JacobianForElementSize=&JacobianForElementSize;")
        w(" functable->numglobal_params=%d;" % len(code.global_params))
        w(" functable->global_parameters=(double **)calloc(%d,sizeof(double*));" % max(1, len(code.global_params)))
        w(" functable->ResidualAndJacobian=(JITFuncSpec_ResidualAndJacobian_FiniteElement *)calloc(%d,sizeof(JITFuncSpec_ResidualAndJacobian_FiniteElement));" % max(1, len(resnames)))
        w(" functable->ResidualAndJacobianSteady=(JITFuncSpec_ResidualAndJacobian_FiniteElement *)calloc(%d,sizeof(JITFuncSpec_ResidualAndJacobian_FiniteElement));" % max(1, len(resnames)))
        w(" functable->ParameterDerivative=(JITFuncSpec_ResidualAndJacobian_FiniteElement **)calloc(%d,sizeof(JITFuncSpec_ResidualAndJacobian_FiniteElement*));" % max(1, len(resnames)))
        for i, rn in enumerate(resnames):
            w(" functable->ResidualAndJacobian[%d]=&ResidualAndJacobian%d;" % (i, i))
            # steady solves run the same routine with zeroed weights / ntstorage 0 (src/elements.cpp:4583-4596)
            w(" functable->ResidualAndJacobianSteady[%d]=&ResidualAndJacobian%d;" % (i, i))
            w(" functable->ParameterDerivative[%d]=(JITFuncSpec_ResidualAndJacobian_FiniteElement *)calloc(%d,sizeof(JITFuncSpec_ResidualAndJacobian_FiniteElement));" % (i, max(1, len(code.global_params))))
            for k, p in enumerate(code.global_params):
                w(" functable->ParameterDerivative[%d][%d]=&dResidual%ddParameter_%s;" % (i, k, i, p))
        w(" functable->max_dt_order=%d;" % code.max_dt_order())
        w(" functable->moving_nodes=%s;" % ("true" if code.coordinates_as_dofs else "false"))
        w(" functable->integration_order=0;")
        w(' SET_INTERNAL_NAME(functable->dominant_space,"C2");')
        w(" functable->hessian_generated=%s;" % ("true" if hess else "false"))
        if hess:
            w(" functable->shapes_required_Hessian=(JITFuncSpec_RequiredShapes_FiniteElement_t *)calloc(%d,sizeof(JITFuncSpec_RequiredShapes_FiniteElement_t));" % max(1, len(resnames)))
            w(" functable->HessianVectorProduct=(JITFuncSpec_HessianVectorProduct_FiniteElement *)calloc(%d,sizeof(JITFuncSpec_HessianVectorProduct_FiniteElement));" % max(1, len(resnames)))
            for i, rn in enumerate(resnames):
                w(" functable->HessianVectorProduct[%d]=&HessianVectorProduct%d;" % (i, i))
        if nint:      # src/codegen.cpp:6991-7003
            w(" functable->numintegral_expressions=%d;" % nint)
            w(" functable->integral_expressions_names=(char **)malloc(sizeof(char*)*functable->numintegral_expressions);")
            for i, n in enumerate(code.integral_expressions.keys()):
                w(' SET_INTERNAL_FIELD_NAME(functable->integral_expressions_names,%d,"%s");' % (i, n))
            w(" functable->EvalIntegralExpression=&EvalIntegralExpression;")
        # local / extremum expressions and Z2 fluxes (src/codegen.cpp:7008-7040, :6786-6796); names are not freed by this test plugin
        if code.local_expressions:
            w(" functable->numlocal_expressions=%d;" % len(code.local_expressions))
            w(" functable->local_expressions_names=(char **)malloc(sizeof(char*)*functable->numlocal_expressions);")
            for i, n in enumerate(code.local_expressions.keys()):
                w(' SET_INTERNAL_FIELD_NAME(functable->local_expressions_names,%d,"%s");' % (i, n))
            w(" functable->EvalLocalExpression=&EvalLocalExpression;")
        if code.extremum_expressions:
            w(" functable->numextremum_expressions=%d;" % len(code.extremum_expressions))
            w(" functable->extremum_expressions_names=(char **)malloc(sizeof(char*)*functable->numextremum_expressions);")
            for i, n in enumerate(code.extremum_expressions.keys()):
                w(' SET_INTERNAL_FIELD_NAME(functable->extremum_expressions_names,%d,"%s");' % (i, n))
            w(" functable->EvalExtremumExpression=&EvalExtremumExpression;")
        if code.Z2_fluxes:
            w(" functable->num_Z2_flux_terms = %d;" % len(code.Z2_fluxes))
            w(" functable->GetZ2Fluxes=&GetZ2Fluxes;")
        w(" functable->clean_up=&clean_up;")
        w(" my_func_table=functable;")
        w("}")
        return "\n".join(o) + "\n"


def emit_plugin_source(code: FiniteElementCode) -> str:
    return CEmitter(code).emit()
