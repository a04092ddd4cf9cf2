// hashCombine / getVariantFromSharedPtr (same names as the reference's VulkanWrapper/utils.hpp:8-19).
#pragma once

#include <functional>
#include <memory>
#include <variant>

template <typename T, typename... Rest>
void hashCombine(std::size_t& seed, const T& v, const Rest&... rest) {
	seed ^= std::hash<T>{}(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
	(hashCombine(seed, rest), ...);
}

template <typename Alt, typename... Types>
auto getVariantFromSharedPtr(std::shared_ptr<std::variant<Types...>> v) -> Alt* {
	return v ? std::get_if<Alt>(v.get()) : nullptr;
}
