// Concept checks are compile-time documentation in ReaK; they do not affect results.
#ifndef RKB_SHIM_BOOST_CONCEPT_CHECK_HPP
#define RKB_SHIM_BOOST_CONCEPT_CHECK_HPP
#define BOOST_CONCEPT_ASSERT(ModelInParens) static_assert(true, "")
#define BOOST_CONCEPT_USAGE(model) void rkb_shim_concept_usage_##model()
namespace boost {
template <typename T> void ignore_unused_variable_warning(const T&) {}
template <typename T> struct CopyConstructibleConcept {};
template <typename T> struct AssignableConcept {};
template <typename T> struct DefaultConstructibleConcept {};
template <typename T> struct EqualityComparableConcept {};
template <typename T> struct LessThanComparableConcept {};
template <typename T> struct CopyConstructible {};
template <typename T> struct Assignable {};
template <typename T> struct DefaultConstructible {};
template <typename T> struct EqualityComparable {};
template <typename T> struct LessThanComparable {};
}
#endif
