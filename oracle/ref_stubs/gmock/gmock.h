// Stand-in for gmock/gmock.h: the three matchers the reference's tests use
// (Each(Field(&T::member, DoubleEq(x)))) and EXPECT_THAT.  TEST INFRASTRUCTURE
// (oracle/Makefile, target ref_gtests; see gtest/gtest.h in this directory).
#pragma once

#include <gtest/gtest.h>

namespace testing {

struct DoubleEqMatcher {
  double want;
  bool operator()(double got) const { return internal::double_eq(want, got); }
};
inline DoubleEqMatcher DoubleEq(double want) { return DoubleEqMatcher{want}; }

template <class C, class M, class Inner> struct FieldMatcher {
  M C::*member;
  Inner inner;
  bool operator()(const C &object) const { return inner(object.*member); }
};
template <class C, class M, class Inner> FieldMatcher<C, M, Inner> Field(M C::*member, Inner inner) {
  return FieldMatcher<C, M, Inner>{member, inner};
}

template <class Inner> struct EachMatcher {
  Inner inner;
  template <class Container> bool operator()(const Container &c) const {
    for (const auto &item : c)
      if (!inner(item)) return false;
    return true;
  }
};
template <class Inner> EachMatcher<Inner> Each(Inner inner) { return EachMatcher<Inner>{inner}; }

}  // namespace testing

#define EXPECT_THAT(value, matcher) QMCL_GTEST_REPORT((matcher)(value), "EXPECT_THAT(" #value ", " #matcher ")")
