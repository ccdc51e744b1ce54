// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's IKFoM toolkit): Boost is not installed in this image.
// This header implements, from scratch, exactly the Boost.Preprocessor macros mtk/build_manifold.hpp uses, for
// sequences of up to 12 elements (state_ikfom has 8): SEQ_SIZE / HEAD / TAIL / ENUM / TRANSFORM_S / FOR_EACH_R, FOR_1,
// IF, DEC and the two old-style tuple helpers.  Same observable expansion as Boost's for these uses.
#pragma once

#define BOOST_PP_CAT(a, b) BOOST_PP_CAT_I(a, b)
#define BOOST_PP_CAT_I(a, b) a##b

// ---- small integer tables
#define LSDPP_DEC_0 0
#define LSDPP_DEC_1 0
#define LSDPP_DEC_2 1
#define LSDPP_DEC_3 2
#define LSDPP_DEC_4 3
#define LSDPP_DEC_5 4
#define LSDPP_DEC_6 5
#define LSDPP_DEC_7 6
#define LSDPP_DEC_8 7
#define LSDPP_DEC_9 8
#define LSDPP_DEC_10 9
#define LSDPP_DEC_11 10
#define LSDPP_DEC_12 11
#define LSDPP_DEC_13 12
#define BOOST_PP_DEC(x) BOOST_PP_CAT(LSDPP_DEC_, x)

#define LSDPP_BOOL_0 0
#define LSDPP_BOOL_1 1
#define LSDPP_BOOL_2 1
#define LSDPP_BOOL_3 1
#define LSDPP_BOOL_4 1
#define LSDPP_BOOL_5 1
#define LSDPP_BOOL_6 1
#define LSDPP_BOOL_7 1
#define LSDPP_BOOL_8 1
#define LSDPP_BOOL_9 1
#define LSDPP_BOOL_10 1
#define LSDPP_BOOL_11 1
#define LSDPP_BOOL_12 1
#define LSDPP_BOOL_13 1
#define BOOST_PP_BOOL(x) BOOST_PP_CAT(LSDPP_BOOL_, x)

#define LSDPP_IIF_0(t, f) f
#define LSDPP_IIF_1(t, f) t
#define BOOST_PP_IIF(bit, t, f) BOOST_PP_CAT(LSDPP_IIF_, bit)(t, f)
#define BOOST_PP_IF(cond, t, f) BOOST_PP_IIF(BOOST_PP_BOOL(cond), t, f)

// ---- old-style tuple helpers (Boost < 1.50 names that build_manifold.hpp still uses)
#define BOOST_PP_TUPLE_REM_2(a, b) a, b
#define BOOST_PP_TUPLE_ELEM_2_0(a, b) a

// ---- sequences  (a)(b)(c)
#define BOOST_PP_SEQ_HEAD(seq) LSDPP_SEQ_HEAD_I(LSDPP_SEQ_HEAD_II seq)
#define LSDPP_SEQ_HEAD_II(x) x, LSDPP_NIL
#define LSDPP_SEQ_HEAD_I(...) LSDPP_FIRST(__VA_ARGS__)
#define LSDPP_FIRST(x, ...) x
#define BOOST_PP_SEQ_TAIL(seq) LSDPP_SEQ_TAIL_I seq
#define LSDPP_SEQ_TAIL_I(x)

#define BOOST_PP_SEQ_SIZE(seq) BOOST_PP_CAT(LSDPP_SEQ_SIZE_, LSDPP_SEQ_SIZE_0 seq)
#define LSDPP_SEQ_SIZE_0(_) LSDPP_SEQ_SIZE_1
#define LSDPP_SEQ_SIZE_1(_) LSDPP_SEQ_SIZE_2
#define LSDPP_SEQ_SIZE_2(_) LSDPP_SEQ_SIZE_3
#define LSDPP_SEQ_SIZE_3(_) LSDPP_SEQ_SIZE_4
#define LSDPP_SEQ_SIZE_4(_) LSDPP_SEQ_SIZE_5
#define LSDPP_SEQ_SIZE_5(_) LSDPP_SEQ_SIZE_6
#define LSDPP_SEQ_SIZE_6(_) LSDPP_SEQ_SIZE_7
#define LSDPP_SEQ_SIZE_7(_) LSDPP_SEQ_SIZE_8
#define LSDPP_SEQ_SIZE_8(_) LSDPP_SEQ_SIZE_9
#define LSDPP_SEQ_SIZE_9(_) LSDPP_SEQ_SIZE_10
#define LSDPP_SEQ_SIZE_10(_) LSDPP_SEQ_SIZE_11
#define LSDPP_SEQ_SIZE_11(_) LSDPP_SEQ_SIZE_12
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_0 0
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_1 1
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_2 2
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_3 3
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_4 4
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_5 5
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_6 6
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_7 7
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_8 8
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_9 9
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_10 10
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_11 11
#define LSDPP_SEQ_SIZE_LSDPP_SEQ_SIZE_12 12

// (a)(b)(c) -> a, b, c
#define BOOST_PP_SEQ_ENUM(seq) BOOST_PP_CAT(LSDPP_SEQ_ENUM_, BOOST_PP_SEQ_SIZE(seq)) seq
#define LSDPP_SEQ_ENUM_1(x) x
#define LSDPP_SEQ_ENUM_2(x) x, LSDPP_SEQ_ENUM_1
#define LSDPP_SEQ_ENUM_3(x) x, LSDPP_SEQ_ENUM_2
#define LSDPP_SEQ_ENUM_4(x) x, LSDPP_SEQ_ENUM_3
#define LSDPP_SEQ_ENUM_5(x) x, LSDPP_SEQ_ENUM_4
#define LSDPP_SEQ_ENUM_6(x) x, LSDPP_SEQ_ENUM_5
#define LSDPP_SEQ_ENUM_7(x) x, LSDPP_SEQ_ENUM_6
#define LSDPP_SEQ_ENUM_8(x) x, LSDPP_SEQ_ENUM_7
#define LSDPP_SEQ_ENUM_9(x) x, LSDPP_SEQ_ENUM_8
#define LSDPP_SEQ_ENUM_10(x) x, LSDPP_SEQ_ENUM_9
#define LSDPP_SEQ_ENUM_11(x) x, LSDPP_SEQ_ENUM_10
#define LSDPP_SEQ_ENUM_12(x) x, LSDPP_SEQ_ENUM_11

// element-wise application: macro(r, data, elem) for every element, results juxtaposed (FOR_EACH) or re-wrapped (TRANSFORM)
#define BOOST_PP_SEQ_FOR_EACH_R(r, macro, data, seq) BOOST_PP_CAT(LSDPP_SFE_, BOOST_PP_SEQ_SIZE(seq))(macro, data, seq)
#define LSDPP_SFE_1(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s))
#define LSDPP_SFE_2(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_1(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_3(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_2(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_4(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_3(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_5(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_4(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_6(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_5(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_7(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_6(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_8(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_7(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_9(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_8(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_10(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_9(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_11(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_10(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_SFE_12(m, d, s) m(2, d, BOOST_PP_SEQ_HEAD(s)) LSDPP_SFE_11(m, d, BOOST_PP_SEQ_TAIL(s))

#define BOOST_PP_SEQ_TRANSFORM_S(s_, op, data, seq) BOOST_PP_CAT(LSDPP_ST_, BOOST_PP_SEQ_SIZE(seq))(op, data, seq)
#define LSDPP_ST_1(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s)))
#define LSDPP_ST_2(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_1(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_3(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_2(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_4(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_3(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_5(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_4(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_6(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_5(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_7(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_6(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_8(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_7(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_9(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_8(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_10(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_9(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_11(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_10(m, d, BOOST_PP_SEQ_TAIL(s))
#define LSDPP_ST_12(m, d, s) (m(2, d, BOOST_PP_SEQ_HEAD(s))) LSDPP_ST_11(m, d, BOOST_PP_SEQ_TAIL(s))

// ---- BOOST_PP_FOR_1(state, pred, op, macro): while pred(r, state): macro(r, state); state = op(r, state)
#define LSDPP_EAT(...)
#define LSDPP_FOR_STEP(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define BOOST_PP_FOR_1(s, p, o, m) LSDPP_FOR_STEP(2, LSDPP_FOR_2, s, p, o, m)
#define LSDPP_FOR_2(s, p, o, m) LSDPP_FOR_STEP_B(3, LSDPP_FOR_3, s, p, o, m)
#define LSDPP_FOR_3(s, p, o, m) LSDPP_FOR_STEP_C(4, LSDPP_FOR_4, s, p, o, m)
#define LSDPP_FOR_4(s, p, o, m) LSDPP_FOR_STEP_D(5, LSDPP_FOR_5, s, p, o, m)
#define LSDPP_FOR_5(s, p, o, m) LSDPP_FOR_STEP_E(6, LSDPP_FOR_6, s, p, o, m)
#define LSDPP_FOR_6(s, p, o, m) LSDPP_FOR_STEP_F(7, LSDPP_FOR_7, s, p, o, m)
#define LSDPP_FOR_7(s, p, o, m) LSDPP_FOR_STEP_G(8, LSDPP_FOR_8, s, p, o, m)
#define LSDPP_FOR_8(s, p, o, m) LSDPP_FOR_STEP_H(9, LSDPP_FOR_9, s, p, o, m)
#define LSDPP_FOR_9(s, p, o, m) LSDPP_FOR_STEP_I(10, LSDPP_FOR_10, s, p, o, m)
#define LSDPP_FOR_10(s, p, o, m) LSDPP_FOR_STEP_J(11, LSDPP_FOR_11, s, p, o, m)
#define LSDPP_FOR_11(s, p, o, m) LSDPP_FOR_STEP_K(12, LSDPP_FOR_12, s, p, o, m)
#define LSDPP_FOR_12(s, p, o, m) LSDPP_FOR_STEP_L(13, LSDPP_FOR_13, s, p, o, m)
#define LSDPP_FOR_13(s, p, o, m)
// one copy of STEP / BODY per depth: a macro cannot re-enter itself while it is being expanded
#define LSDPP_FOR_STEP_B(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_B, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_B(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_C(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_C, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_C(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_D(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_D, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_D(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_E(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_E, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_E(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_F(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_F, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_F(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_G(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_G, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_G(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_H(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_H, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_H(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_I(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_I, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_I(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_J(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_J, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_J(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_K(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_K, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_K(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
#define LSDPP_FOR_STEP_L(n, next, s, p, o, m) BOOST_PP_IIF(BOOST_PP_BOOL(p(n, s)), LSDPP_FOR_BODY_L, LSDPP_EAT)(n, next, s, p, o, m)
#define LSDPP_FOR_BODY_L(n, next, s, p, o, m) m(n, s) next(o(n, s), p, o, m)
