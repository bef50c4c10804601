#pragma once
namespace OpenMM {
class Context {};
}
