#pragma once
// stand-in: cvo.hpp includes it, nothing on the path uses it
